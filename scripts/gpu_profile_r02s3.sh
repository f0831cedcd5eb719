#!/bin/bash
# ncu evidence for the final bench of round 2 (one gpurun call; each ncu run only after the same command exited 0 plainly).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/prof_plain2.json 2> gpurun_out/prof_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:rollout_e32_kernel -s 4 -c 1 -f -o gpurun_out/prof_rollout $CMD > gpurun_out/ncu_full_rollout.log 2>&1
echo "full capture (rollout) rc=$?"
python scripts/time_recover.py 16842752 > gpurun_out/recover_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_recover_umma -s 2 -c 1 -f -o gpurun_out/prof_recover python scripts/time_recover.py 16842752 > gpurun_out/ncu_full_recover.log 2>&1
echo "full capture (recover) rc=$?"
python scripts/time_forward.py > gpurun_out/forward_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:forward_e32_tc -s 2 -c 1 -f -o gpurun_out/prof_forward python scripts/time_forward.py > gpurun_out/ncu_full_forward.log 2>&1
echo "full capture (forward) rc=$?"
ls -la gpurun_out
