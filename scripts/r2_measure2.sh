#!/bin/bash
# Round-2 second pass (run under gpurun): GPU parity tests, kernel variants (build_variants/*.so, LB_DECODE modes) at a
# mid-size batch, launch lists and `ncu --set full` captures summarised ON THE BOX.  Every command reads /dev/null and
# is bounded by `timeout -k` (a script that waited on stdin cost the first pass its ncu half).
O=gpurun_out
mkdir -p $O
T=${TAG:-r2}
exec < /dev/null
TO="timeout -k 10"
bench() {  # name [env...] -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" $TO 240 python bench.py "$@" --no-e2e --cpu-sample-docs 4 > $O/${T}_var_$name.json 2> $O/var_$name.err
}
cap() {  # name regex skip config docs [env...]
  local name=$1 rx=$2 skip=$3 cfg=$4 docs=$5; shift 5
  env "$@" $TO 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $O/prof_${name}_$T -f \
      python bench.py --config $cfg --docs $docs --steps 1 --warmup 1 --no-e2e --cpu-sample-docs 4 > /dev/null 2>$O/prof_$name.err
  $TO 120 python scripts/summarize_one.py $O/prof_${name}_$T.ncu-rep $name > $O/${T}_ncu_$name.md 2>$O/sum_$name.err
  rm -f $O/prof_${name}_$T.ncu-rep
}
date > $O/${T}_pass2_times.txt
$TO 600 python -m pytest tests -m gpu -x -q > $O/${T}_gputests.log 2>&1; tail -3 $O/${T}_gputests.log
date >> $O/${T}_pass2_times.txt
# ---- variants at 8192 documents of C3 (default build = leaf prefetch for deletes and inserts)
bench C3_default -- --docs 8192 --steps 3 --warmup 3
for v in build_variants/*.so; do
  n=$(basename $v .so)
  bench C3_$n LORO_B200_LIB=$PWD/$v -- --docs 8192 --steps 3 --warmup 3
done
bench C3_group LB_DECODE=group -- --docs 8192 --steps 3 --warmup 3
bench C3_warp LB_DECODE=warp -- --docs 8192 --steps 3 --warmup 3
bench C2_default -- --config C2 --docs 512 --steps 3 --warmup 3
bench C2_pf0 LORO_B200_LIB=$PWD/build_variants/pf0.so -- --config C2 --docs 512 --steps 3 --warmup 3
bench C2_group LB_DECODE=group -- --config C2 --docs 512 --steps 3 --warmup 3
bench C5_default -- --config C5 --docs 4096 --steps 3 --warmup 3
bench C5_group LB_DECODE=group -- --config C5 --docs 4096 --steps 3 --warmup 3
bench C4_default -- --config C4 --c4-base 200000 --c4-peers 64 --c4-edits 5000 --steps 2 --warmup 1
bench C4_pf0 LORO_B200_LIB=$PWD/build_variants/pf0.so -- --config C4 --c4-base 200000 --c4-peers 64 --c4-edits 5000 --steps 2 --warmup 1
python scripts/show_bench.py $O/${T}_var_*.json > $O/${T}_variants.txt 2>&1
cat $O/${T}_variants.txt
date >> $O/${T}_pass2_times.txt
# ---- launch lists
$TO 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_C3.csv python bench.py --docs 8192 --steps 1 --warmup 1 --no-e2e --cpu-sample-docs 4 > /dev/null 2>&1
$TO 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_C5.csv python bench.py --config C5 --docs 4096 --steps 1 --warmup 1 --no-e2e --cpu-sample-docs 4 > /dev/null 2>&1
$TO 60 python scripts/launch_summary.py $O/${T}_launches_C3.csv $O/${T}_launches_C5.csv > $O/${T}_launches_summary.md 2>&1
date >> $O/${T}_pass2_times.txt
# ---- full captures (one kernel launch each, summarised here)
cap seq k_seq_integrate 1 C3 2048
cap decode k_block_decode_cols 1 C3 2048
cap decode_group k_block_decode_group 1 C3 2048 LB_DECODE=group
cap expenc k_exp_encode 3 C3 2048
cap expchg k_exp_changes 1 C3 2048
cap tree_apply k_tree_apply 1 C5 2048
date >> $O/${T}_pass2_times.txt
ls -la $O | tail -50
