#!/bin/bash
# last check of the round (run under gpurun): whole GPU suite on the final build, the C5 and C2 lines, a mid-size C3 run
O=gpurun_out
mkdir -p $O
T=r2i
exec < /dev/null
TO="timeout -k 10"
$TO 600 python -m pytest tests -m gpu -x -q > $O/${T}_gputests.log 2>&1; tail -2 $O/${T}_gputests.log
$TO 300 python bench.py --config C5 --steps 5 --warmup 3 > $O/${T}_bench_C5.json 2> $O/bench_C5.err
$TO 300 python bench.py --config C2 --steps 5 --warmup 3 > $O/${T}_bench_C2.json 2> $O/bench_C2.err
$TO 240 python bench.py --docs 8192 --steps 3 --warmup 3 --cpu-sample-docs 4 > $O/${T}_bench_C3_8192.json 2> $O/bench_C3_8192.err
python scripts/show_bench.py $O/${T}_bench_*.json
