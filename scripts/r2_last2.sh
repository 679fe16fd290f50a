#!/bin/bash
# the C2 line on the final build (encoder threshold moved), plus the docset measurement again (documents now keep every blob)
O=gpurun_out
mkdir -p $O
exec < /dev/null
TO="timeout -k 10"
$TO 300 python bench.py --config C2 --steps 5 --warmup 3 > $O/r2j_bench_C2.json 2> $O/bench_C2.err
$TO 300 python scripts/bench_docset.py --docs 8192 --steps 3 > $O/r2j_bench_docset.json 2> $O/bench_docset.err
$TO 200 python -m pytest tests -m gpu -x -q -k "export or docset" > $O/r2j_gputests.log 2>&1; tail -2 $O/r2j_gputests.log
python scripts/show_bench.py $O/r2j_bench_C2.json; cat $O/r2j_bench_docset.json
