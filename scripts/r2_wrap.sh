#!/bin/bash
# end-of-round check (run under gpurun): the whole GPU suite on the final build, smoke(), and the C2 lines
O=gpurun_out
mkdir -p $O
exec < /dev/null
TO="timeout -k 10"
$TO 400 python -m pytest tests -m gpu -x -q > $O/r2z_gputests.log 2>&1; tail -2 $O/r2z_gputests.log
$TO 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z_smoke.log 2>&1; tail -1 $O/r2z_smoke.log | cut -c1-300
$TO 200 python bench.py --config C2 --steps 5 --warmup 3 > $O/r2z_bench_C2.json 2> $O/bench_C2.err
$TO 200 python bench.py --config C2 --c2-distinct-peers --steps 5 --warmup 3 > $O/r2z_bench_C2_distinct_peers.json 2> $O/bench_C2d.err
python scripts/show_bench.py $O/r2z_bench_C2*.json
tail -2 $O/bench_C2d.err
