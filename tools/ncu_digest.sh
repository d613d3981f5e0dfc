#!/bin/bash
# usage: tools/ncu_digest.sh <name> [launch index]   (reads gpurun_out/<name>.ncu-rep)
n=$1; idx=${2:-0}
ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/$n.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/$n.csv | grep -v "pipe_\|occupancy_limit_w\|l1tex__data\|launch__block\|launch__grid"
python tools/ncu_stalls.py raw gpurun_out/$n.csv $idx
ncu -i gpurun_out/$n.ncu-rep --page source --csv > gpurun_out/${n}_src.csv 2>/dev/null
