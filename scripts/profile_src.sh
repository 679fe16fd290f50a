#!/usr/bin/env bash
# Cheap source-level pass (instruction + stall attribution only) for one kernel of the bench command.
set -u
TAG=${1:-x}; DOCS=${2:-4096}; K=${3:-k_seq_integrate}; SKIP=${4:-3}
BENCH="python bench.py --docs $DOCS --steps 2 --warmup 3 --no-e2e --cpu-sample-docs 16"
mkdir -p gpurun_out
ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section InstructionStats --clock-control none --import-source on \
    -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/src_${K}_$TAG $BENCH > gpurun_out/ncu_src_${K}_$TAG.log 2>&1
tail -3 gpurun_out/ncu_src_${K}_$TAG.log
