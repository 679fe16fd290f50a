#!/bin/sh
# round-2 captures (run under gpurun): launch lists + one `ncu --set full` capture per hot kernel.
# C3 at 1024 docs and C5 at 512 docs keep the replays short; shares, not absolutes, are what the launch lists are for.
set -x
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2_C3.csv python bench.py --docs 8192 --steps 1 --warmup 1 --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2_C5.csv python bench.py --config C5 --docs 4096 --steps 1 --warmup 1 --no-e2e > /dev/null 2>&1
cap() {  # name regex skip config docs
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $O/prof_$1_r2 -f python bench.py --config $4 --docs $5 --steps 1 --warmup 1 --no-e2e > /dev/null 2>$O/prof_$1.err
}
cap decode k_block_decode_cols 1 C3 1024
cap seq k_seq_integrate 1 C3 1024
cap expenc k_exp_encode 3 C3 1024
cap expchg k_exp_changes 1 C3 1024
cap tree k_tree_build 1 C5 512
cap json k_json 3 C5 512
ls -la $O/*_r2.ncu-rep
