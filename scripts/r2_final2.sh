#!/bin/bash
# Round-2 last pass (run under gpurun): the encoder variants compared AT FULL SIZE (at 8192 documents the register-capped
# encoder won, at 100 000 it lost), then the final lines of C3 / C5 / C2 with the winner and the split host path (e2e).
O=gpurun_out
mkdir -p $O
T=r2g
exec < /dev/null
TO="timeout -k 10"
date > $O/${T}_times.txt
$TO 600 python -m pytest tests -m gpu -x -q > $O/${T}_gputests.log 2>&1; tail -3 $O/${T}_gputests.log
$TO 300 python bench.py --steps 3 --warmup 3 --no-e2e --cpu-sample-docs 4 > $O/${T}_ab_default.json 2> $O/ab_default.err
for n in xla0 xenc5; do
  LORO_B200_LIB=$PWD/build_variants/$n.so $TO 300 python bench.py --steps 3 --warmup 3 --no-e2e --cpu-sample-docs 4 > $O/${T}_ab_$n.json 2> $O/ab_$n.err
done
USE=$(python scripts/pick_variant.py $T)
echo "final lines use: ${USE:-default build}" | tee $O/${T}_choice.txt
[ -n "$USE" ] && export LORO_B200_LIB=$PWD/build_variants/$USE.so
python scripts/show_bench.py $O/${T}_ab_*.json
date >> $O/${T}_times.txt
$TO 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench_C3.json 2> $O/bench_C3.err
$TO 300 python bench.py --config C5 --steps 5 --warmup 3 > $O/${T}_bench_C5.json 2> $O/bench_C5.err
$TO 300 python bench.py --config C2 --steps 5 --warmup 3 > $O/${T}_bench_C2.json 2> $O/bench_C2.err
python scripts/show_bench.py $O/${T}_bench_C*.json
date >> $O/${T}_times.txt
ls -la $O | tail -20
