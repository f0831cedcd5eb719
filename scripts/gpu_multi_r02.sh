#!/bin/bash
# Round-2 multi-GPU evidence on N GPUs of one box: the bench line (weak value + strong record + training step with the
# all-reduce), the Koopman and transformer training steps at N/2 and N ranks.
N=${1:-8}
H=$((N / 2))
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n$N.err
for R in $H $N; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port 29521 scripts/bench_train.py > gpurun_out/r02_train_n$R.jsonl 2> gpurun_out/r02_train_n$R.err
  echo "train n=$R rc=$?"; cat gpurun_out/r02_train_n$R.jsonl
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port 29531 scripts/bench_train_transformer.py --batches 16,148 > gpurun_out/r02_train_tf_n$R.jsonl 2> gpurun_out/r02_train_tf_n$R.err
  echo "train_tf n=$R rc=$?"; cat gpurun_out/r02_train_tf_n$R.jsonl
done
