#!/bin/bash
# usage: gpu_ncu.sh <outname> <python script args...>   (ncu --set full of one rollout kernel launch)
OUT=$1; shift
mkdir -p gpurun_out
python scripts/prof_rollout.py "$@" > gpurun_out/${OUT}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout -s 1 -c 1 -f -o gpurun_out/$OUT python scripts/prof_rollout.py "$@" > gpurun_out/${OUT}_ncu.log 2>&1
echo "rc=$?"; cat gpurun_out/${OUT}_plain.log; tail -2 gpurun_out/${OUT}_ncu.log
