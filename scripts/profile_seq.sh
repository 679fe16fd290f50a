#!/usr/bin/env bash
# Short profiling pass: `--set full` capture of k_seq_integrate (+ optionally another kernel) for the bench command.
set -u
TAG=${1:-x}
DOCS=${2:-8192}
K2=${3:-}
BENCH="python bench.py --docs $DOCS --steps 2 --warmup 3 --no-e2e --cpu-sample-docs 16"
mkdir -p gpurun_out
$BENCH > gpurun_out/bench_${TAG}_profcfg.json 2> gpurun_out/bench_${TAG}_profcfg.err
ncu --set full --clock-control none --import-source on -k regex:k_seq_integrate -s 3 -c 1 -f -o gpurun_out/prof_seq_$TAG \
    $BENCH > gpurun_out/ncu_seq_$TAG.log 2>&1
if [ -n "$K2" ]; then
ncu --set full --clock-control none --import-source on -k regex:$K2 -s 6 -c 1 -f -o gpurun_out/prof_${K2}_$TAG \
    $BENCH > gpurun_out/ncu_${K2}_$TAG.log 2>&1
fi
