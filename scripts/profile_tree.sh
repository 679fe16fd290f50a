#!/bin/bash
# tree-path captures (run under gpurun): launch list of config C5 + one `ncu --set full` capture per tree kernel, summarised ON THE BOX
O=gpurun_out
mkdir -p $O
cap() {  # name regex skip docs
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $O/prof_$1_r2 -f python bench.py --config C5 --docs $4 --steps 1 --warmup 1 --no-e2e > /dev/null 2>$O/prof_$1.err
  python scripts/summarize_one.py $O/prof_$1_r2.ncu-rep $1 > $O/r2_ncu_$1.md 2>$O/sum_$1.err
  rm -f $O/prof_$1_r2.ncu-rep
}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2_C5.csv python bench.py --config C5 --steps 1 --warmup 1 --no-e2e > /dev/null 2>&1
cap tree_apply k_tree_apply 1 4096
cap tree_layout k_tree_layout 1 4096
cap json_tree k_json 3 4096
