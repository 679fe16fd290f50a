#!/bin/bash
# two GPUs of one box: the C4 digests test (one GPU), then the 2-rank bench line (torchrun, NCCL counters all-gather)
O=gpurun_out
mkdir -p $O
exec < /dev/null
TO="timeout -k 10"
$TO 200 python -m pytest tests -m gpu -x -q -k "c4_full_size" > $O/r2k_gputests_c4.log 2>&1; tail -2 $O/r2k_gputests_c4.log
$TO 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --docs 16384 --steps 3 --warmup 3 --cpu-sample-docs 4 > $O/r2k_bench_2gpu_16k.json 2> $O/bench_2gpu.err
tail -1 $O/r2k_bench_2gpu_16k.json | cut -c1-600
tail -3 $O/bench_2gpu.err
