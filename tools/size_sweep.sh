#!/bin/bash
# two-kernel path vs every fused-step variant, per size / dtype
for cfg in "256 64 2048 2" "512 64 512 3" "1024 64 128 7" "2048 64 32 2" "256 128 1024 2" "512 128 256 2" "1024 128 64 4" "2048 128 16 1"; do
  set -- $cfg
  LP2D_PIPE=0 python tools/pipe_sweep.py $1 $2 $3 "PIPE=0" 2>&1 | tail -1
  for ((v=0; v<$4; v++)); do python tools/pipe_sweep.py $1 $2 $3 "PIPE=2;OPS=7;VAR=$v" 2>&1 | tail -1; done
done
