#!/bin/bash
# round 2, first GPU call: A/B of the two prepared experiments + balanced tile sizes (MGB_DEBUG_E) per round
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
QUICK=1 bash scripts/gpu_shot_ab.sh winv ow both > gpurun_out/r2a_ab.txt 2>&1
for lib in "" _winv; do
  export MGB_LIB=$PWD/montgomery_b200/libmontgomery_b200$lib.so
  for cfg in "64,64,32,16,8" "56,56,28,14,7" "56,56,28,14,8" "37,56,28,14,7" "56,28,28,14,7" "56,56,28,28,14"; do
    echo "lib=$lib E=$cfg"
    MGB_DEBUG_NBIG=4 MGB_DEBUG_E=$cfg MGB_DEBUG_ROUNDS=1 timeout 60 python scripts/quick_time.py 20 2>&1 | grep -E "round|accumulate" | tail -6 | sed -E "s/.*('accumulate': [0-9.]+).*('total': [0-9.]+).*/\1 \2/"
  done
done > gpurun_out/r2a_esweep.txt 2>&1
cat gpurun_out/r2a_ab.txt | tail -8
cat gpurun_out/r2a_esweep.txt
