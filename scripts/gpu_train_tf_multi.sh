#!/bin/bash
# data-parallel transformer training step on N GPUs (one process per GPU, NCCL all-reduce of the gradient blob)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/bench_train_transformer.py --batches 16,148 > gpurun_out/train_tf_n$N.jsonl 2> gpurun_out/train_tf_n$N.err
echo "rc=$?"; cat gpurun_out/train_tf_n$N.jsonl; tail -3 gpurun_out/train_tf_n$N.err
