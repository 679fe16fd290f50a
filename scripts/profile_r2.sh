#!/bin/sh
# round-2 captures (run under gpurun): launch lists + one `ncu --set full` capture per hot kernel, summarised ON THE BOX
# (scripts/summarize_one.py) because gpurun brings back at most 64 MiB and a capture with sources is ~25 MB.
# usage: profile_r2.sh [names...]   names from: lists decode seq expenc expchg tree json
O=gpurun_out
want() { [ $# -eq 0 ] && return 0; }
cap() {  # name regex skip config docs
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $O/prof_$1_r2 -f python bench.py --config $4 --docs $5 --steps 1 --warmup 1 --no-e2e > /dev/null 2>$O/prof_$1.err
  python scripts/summarize_one.py $O/prof_$1_r2.ncu-rep $1 > $O/r2_ncu_$1.md 2>$O/sum_$1.err
  rm -f $O/prof_$1_r2.ncu-rep
}
for n in "$@"; do
  case $n in
    lists)
      ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2_C3.csv python bench.py --docs 8192 --steps 1 --warmup 1 --no-e2e > /dev/null 2>&1
      ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2_C5.csv python bench.py --config C5 --docs 4096 --steps 1 --warmup 1 --no-e2e > /dev/null 2>&1 ;;
    decode) cap decode k_block_decode_cols 1 C3 1024 ;;
    seq) cap seq k_seq_integrate 1 C3 1024 ;;
    expenc) cap expenc k_exp_encode 3 C3 1024 ;;
    expchg) cap expchg k_exp_changes 1 C3 1024 ;;
    tree) cap tree k_tree_build 1 C5 512 ;;
    json) cap json k_json 3 C5 512 ;;
  esac
done
ls -la $O | tail -20
