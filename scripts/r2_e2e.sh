#!/bin/bash
# Host path experiments (run under gpurun): e2e of C3 / C5 / C2 for several numbers of overlapping sub-batches
# (LORO_B200_SPLIT), then the full C3 line with the best one.
O=gpurun_out
mkdir -p $O
T=r2h
exec < /dev/null
TO="timeout -k 10"
date > $O/${T}_times.txt
$TO 300 python -m pytest tests -m gpu -x -q -k "split or docset or small_mixed" > $O/${T}_gputests.log 2>&1; tail -2 $O/${T}_gputests.log
for S in 1 2 4 8; do
  LORO_B200_SPLIT=$S $TO 300 python bench.py --steps 3 --warmup 1 --cpu-sample-docs 4 > $O/${T}_e2e_C3_s$S.json 2> $O/e2e_C3_s$S.err
done
for S in 1 2 4; do
  LORO_B200_SPLIT=$S $TO 200 python bench.py --config C5 --steps 3 --warmup 1 --cpu-sample-docs 4 > $O/${T}_e2e_C5_s$S.json 2> $O/e2e_C5_s$S.err
  LORO_B200_SPLIT=$S $TO 200 python bench.py --config C2 --steps 3 --warmup 1 --cpu-sample-docs 4 > $O/${T}_e2e_C2_s$S.json 2> $O/e2e_C2_s$S.err
done
python scripts/show_bench.py $O/${T}_e2e_*.json | sed 's/{.*}//' 
date >> $O/${T}_times.txt
BEST=$(python - <<'PY'
import json
best, bv = 4, 0
for s in (1, 2, 4, 8):
    try:
        v = json.loads(open(f"gpurun_out/r2h_e2e_C3_s{s}.json").read().strip().splitlines()[-1])["e2e"]["value"]
        if v > bv: best, bv = s, v
    except Exception:
        pass
print(best)
PY
)
echo "C3 final line with LORO_B200_SPLIT=$BEST" | tee $O/${T}_choice.txt
LORO_B200_SPLIT=$BEST $TO 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench_C3.json 2> $O/bench_C3.err
python scripts/show_bench.py $O/${T}_bench_C3.json | sed 's/{.*}//'
date >> $O/${T}_times.txt
