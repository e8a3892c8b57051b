#!/bin/bash
set -u
qt() { timeout 60 python scripts/quick_time.py "$@" 2>&1 | tail -1 | sed -E "s/.*('accumulate': [0-9.]+).*('total': [0-9.]+).*/\1 \2/"; }
echo -n "default: "; qt 20; echo -n "default: "; qt 20
for s in 4 3 2 1 0; do echo -n "STAGGER=$s: "; MGB_DEBUG_STAGGER=$s qt 20; done
echo "per-round, STAGGER=2:"; MGB_DEBUG_STAGGER=2 MGB_DEBUG_ROUNDS=1 timeout 60 python scripts/quick_time.py 20 2>&1 | grep round | tail -5
echo "per-round, default:"; MGB_DEBUG_ROUNDS=1 timeout 60 python scripts/quick_time.py 20 2>&1 | grep round | tail -5
for s in 3 2 1 0; do echo -n "2^16 STAGGER=$s: "; MGB_DEBUG_STAGGER=$s qt 16; done; echo -n "2^16 default: "; qt 16
for s in 2 1 0; do echo -n "2^22 STAGGER=$s: "; MGB_DEBUG_STAGGER=$s qt 22; done; echo -n "2^22 default: "; qt 22
