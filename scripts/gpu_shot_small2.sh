#!/bin/bash
set -u
qt() { timeout 60 python scripts/quick_time.py "$@" 2>&1 | tail -1 | sed -E "s/.*('accumulate': [0-9.]+).*('reduce': [0-9.]+).*('final_sum': [0-9.]+).*('total': [0-9.]+).*('c': [0-9]+).*('K': [0-9]+).*('rounds': [0-9]+).*/\1 \2 \3 \4 \5 \6 \7/"; }
for c in 12 13 14; do for nr in 1 2 3 4; do echo -n "bls 2^16 c=$c NROUNDS=$nr: "; MGB_DEBUG_NROUNDS=$nr qt 16 bls12-377 $c; done; done
for c in 15 16; do for nr in 2 3 4; do echo -n "pallas 2^18 c=$c NROUNDS=$nr: "; MGB_DEBUG_NROUNDS=$nr qt 18 pallas $c; done; done
for c in 13 14; do for nr in 3 4 5 6; do echo -n "ed 2^18 c=$c NROUNDS=$nr: "; MGB_DEBUG_NROUNDS=$nr qt 18 ed-on-bls12-377 $c; done; done
for c in 15 16; do for nr in 2 3 4; do echo -n "bls 2^18 c=$c NROUNDS=$nr: "; MGB_DEBUG_NROUNDS=$nr qt 18 bls12-377 $c; done; done
