#!/bin/bash
# ncu launch list (durations + DRAM bytes per launch) of one eager MFT train step (tools/ncu_step.py) and its per-kernel summary.
# Usage (GPU box): bash tools/ncu_launch_list.sh <tag>      -> gpurun_out/<tag>_launches.csv, <tag>_launch_summary.txt
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
python tools/ncu_step.py > $OUT/${TAG}_step_plain.log 2>&1 || { echo "plain step failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --page raw --log-file $OUT/${TAG}_launches.csv python tools/ncu_step.py > $OUT/${TAG}_ncu_list.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_summary.txt 2>&1
head -25 $OUT/${TAG}_launch_summary.txt
