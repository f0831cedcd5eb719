#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, tuning sweep.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -40 gpurun_out/pytest.log
echo "== bench" ; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; cat gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
echo "== sweep" ; timeout 900 python scripts/sweep_rollout.py --kinds rossler,lorenz --batches 65536 --warps 8,11 --groups 8,4 --variants 0 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err ; echo "sweep rc=$?" ; tail -5 gpurun_out/sweep.err
