# bench every library under build_variants/ (kernel tuning experiments) at a mid-size batch
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in build_variants/*.so; do
  v=$(basename $lib .so)
  LORO_B200_LIB=$PWD/$lib python bench.py --docs ${DOCS:-8192} --steps 2 --warmup 2 --no-e2e --cpu-sample-docs 16 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/var_$v.json').read().strip().splitlines()[-1])
print('$v', round(d['value']/1e6,1), 'Mops/s', {k: round(x,1) for k,x in d['phases_ms'].items()})"
done
