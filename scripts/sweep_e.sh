#!/bin/bash
# tuning aid: per-round time of the accumulation for different tile sizes E
for cfg in "32,32,32,32,32,32" "32,16,16,16,16,16" "16,16,16,16,16,16" "16,8,8,8,8,8" "8,8,8,8,8,8" "32,16,8,4,4,4" "4,4,4,4,4,4"; do
  echo "E=$cfg"
  MGB_DEBUG_E=$cfg MGB_DEBUG_ROUNDS=1 python scripts/quick_time.py 20 2>&1 | tail -8 | grep -E "round|accumulate" | tail -7 | sed -E 's/.*(accumulate.: [0-9.]+).*/\1/'
done
