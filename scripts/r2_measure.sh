#!/bin/bash
# Round-2 evidence pass (run under gpurun): GPU parity tests, one bench line per BASELINE config, launch lists and one
# `ncu --set full` capture per hot kernel, summarised ON THE BOX (scripts/summarize_one.py) so that only text travels back.
# usage: r2_measure.sh [tests] [bench] [lists] [ncu] [ref]     (no argument = everything)
O=gpurun_out
mkdir -p $O
T=${TAG:-r2}
all="tests bench lists ncu ref"
[ $# -gt 0 ] && all="$*"
cap() {  # name regex skip config docs
  timeout 420 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $O/prof_$1_$T -f \
      python bench.py --config $4 --docs $5 --steps 1 --warmup 1 --no-e2e --cpu-sample-docs 4 > /dev/null 2>$O/prof_$1.err
  python scripts/summarize_one.py $O/prof_$1_$T.ncu-rep $1 > $O/${T}_ncu_$1.md 2>$O/sum_$1.err
  rm -f $O/prof_$1_$T.ncu-rep
}
for n in $all; do
  case $n in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_gputests.log 2>&1; tail -3 $O/${T}_gputests.log ;;
    bench)
      timeout 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench_C3.json 2> $O/bench_C3.err
      timeout 500 python bench.py --config C5 --steps 5 --warmup 3 > $O/${T}_bench_C5.json 2> $O/bench_C5.err
      timeout 400 python bench.py --config C2 --steps 5 --warmup 3 > $O/${T}_bench_C2.json 2> $O/bench_C2.err
      timeout 500 python bench.py --config C4 --steps 3 --warmup 3 > $O/${T}_bench_C4.json 2> $O/bench_C4.err
      python scripts/show_bench.py $O/${T}_bench_C*.json ;;
    ref)
      timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference_arm.json 2> $O/bench_ref.err ;;
    lists)
      timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_C3.csv python bench.py --docs 8192 --steps 1 --warmup 1 --no-e2e --cpu-sample-docs 4 > $O/${T}_bench_C3_profcfg_under_ncu.json 2>&1
      timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_C5.csv python bench.py --config C5 --docs 4096 --steps 1 --warmup 1 --no-e2e --cpu-sample-docs 4 > /dev/null 2>&1
      timeout 300 python bench.py --docs 8192 --steps 3 --warmup 3 --no-e2e --cpu-sample-docs 4 > $O/${T}_bench_C3_profcfg.json 2>/dev/null
      python scripts/launch_summary.py $O/${T}_launches_C3.csv $O/${T}_launches_C5.csv > $O/${T}_launches_summary.md ;;
    ncu)
      cap decode k_block_decode_cols 1 C3 2048
      cap seq k_seq_integrate 1 C3 2048
      cap expenc k_exp_encode 3 C3 2048
      cap expchg k_exp_changes 1 C3 2048
      cap tree_apply k_tree_apply 1 C5 2048
      cap json k_json 3 C3 2048 ;;
  esac
done
ls -la $O | tail -40
