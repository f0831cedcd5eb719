#!/bin/bash
# usage: gpu_ncu_cluster.sh <outname> <batch> <horizon>   (ncu --set full of one cluster-kernel rollout, Cylinder size)
OUT=$1; B=$2; S=$3
mkdir -p gpurun_out
python scripts/prof_rollout.py --kind cylinder --batch $B --horizon $S --variant 7 > gpurun_out/${OUT}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout -s 1 -c 1 -f -o gpurun_out/$OUT python scripts/prof_rollout.py --kind cylinder --batch $B --horizon $S --variant 7 > gpurun_out/${OUT}_ncu.log 2>&1
echo "rc=$?"; cat gpurun_out/${OUT}_plain.log; tail -2 gpurun_out/${OUT}_ncu.log
