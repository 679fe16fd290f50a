#!/usr/bin/env bash
# Profiling recipe for one round (run on the GPU box through gpurun).  Usage: scripts/profile.sh <tag> [docs]
# Produces, under gpurun_out/: the launch list of the bench command, one `--set full` capture of the dominant
# kernel (k_seq_integrate) and one of the decode kernel, plus the un-profiled bench line of the same command.
set -u
TAG=${1:-r1}
DOCS=${2:-8192}
BENCH="python bench.py --docs $DOCS --steps 2 --warmup 3 --no-e2e --cpu-sample-docs 16"
mkdir -p gpurun_out
echo "nproc=$(nproc) cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null) affinity=$(python -c 'import os;print(len(os.sched_getaffinity(0)))')" > gpurun_out/host_$TAG.txt
$BENCH > gpurun_out/bench_${TAG}_profcfg.json 2> gpurun_out/bench_${TAG}_profcfg.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    $BENCH > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_seq_integrate -s 3 -c 1 -f -o gpurun_out/prof_seq_$TAG \
    $BENCH > gpurun_out/ncu_seq_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_block_decode -s 3 -c 1 -f -o gpurun_out/prof_decode_$TAG \
    $BENCH > gpurun_out/ncu_decode_$TAG.log 2>&1
ls -la gpurun_out
