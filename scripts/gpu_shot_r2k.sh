#!/bin/bash
# round 2, GPU call 11: buckets per partial sum (CH) of the reduction against the window size
set -u
mkdir -p gpurun_out
qt() { timeout 60 python scripts/quick_time.py "$@" 2>&1 | tail -1 | sed -E "s/.*('accumulate': [0-9.]+).*('reduce': [0-9.]+).*('final_sum': [0-9.]+).*('total': [0-9.]+).*('c': [0-9]+).*('K': [0-9]+).*('rounds': [0-9]+).*/\1 \2 \3 \4 \5 \6 \7/"; }
{
for c in 12 13 14 15 16 17 18; do for ch in 1 2 4 8 16; do echo "== bls12-377 2^18 c=$c CH=$ch"; MGB_DEBUG_CH=$ch qt 18 bls12-377 $c; done; done
for c in 12 13 14; do for ch in 1 2 4 8; do echo "== bls12-377 2^16 c=$c CH=$ch"; MGB_DEBUG_CH=$ch qt 16 bls12-377 $c; done; done
for c in 18 19; do for ch in 4 8 16; do echo "== bls12-377 2^22 c=$c CH=$ch"; MGB_DEBUG_CH=$ch qt 22 bls12-377 $c; done; done
} > gpurun_out/r2k_ch.txt 2>&1
cat gpurun_out/r2k_ch.txt
