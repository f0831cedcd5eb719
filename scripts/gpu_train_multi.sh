#!/bin/bash
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/bench_train.py > gpurun_out/train_n$N.jsonl 2> gpurun_out/train_n$N.err
echo "rc=$?"; cat gpurun_out/train_n$N.jsonl; tail -3 gpurun_out/train_n$N.err
