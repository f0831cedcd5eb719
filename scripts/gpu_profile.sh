#!/bin/bash
# ncu evidence for the bench command (one gpurun call; each ncu run only after the same command exited 0 plainly).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu ${BENCH_ARGS}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL_REGEX:-rollout} -s 4 -c 1 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
