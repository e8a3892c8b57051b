#!/bin/bash
# tuning aid: per-round time of the accumulation for different tile shapes (negative = block-level tile)
for cfg in "$@"; do
  echo "E=$cfg"
  MGB_DEBUG_E=$cfg MGB_DEBUG_ROUNDS=1 python scripts/quick_time.py ${LOGN:-20} 2>&1 | grep -E "round|accumulate" | tail -7 | sed -E "s/.*(accumulate.: [0-9.]+).*/\1/"
done
