#!/bin/bash
# round 2, GPU call 10: window sweeps at the sizes between the measured ones; occupancy variants of the reduction kernels
set -u
mkdir -p gpurun_out
qt() { timeout 60 python scripts/quick_time.py "$@" 2>&1 | tail -1 | sed -E "s/.*('decompose_slice': [0-9.]+).*('sort': [0-9.]+).*('accumulate': [0-9.]+).*('reduce': [0-9.]+).*('final_sum': [0-9.]+).*('total': [0-9.]+).*('c': [0-9]+).*('K': [0-9]+).*('rounds': [0-9]+).*/\1 \2 \3 \4 \5 \6 \7 \8 \9/"; }
{
for v in red3 red4; do echo "== lib_$v 2^20"; MGB_LIB=$PWD/montgomery_b200/libmontgomery_b200_$v.so qt 20; MGB_LIB=$PWD/montgomery_b200/libmontgomery_b200_$v.so qt 16; done
for c in 11 12 13; do echo "== bls12-377 2^15 c=$c"; qt 15 bls12-377 $c; done
for c in 13 14 15 16; do echo "== bls12-377 2^17 c=$c"; qt 17 bls12-377 $c; done
for c in 15 16 17 18; do echo "== bls12-377 2^19 c=$c"; qt 19 bls12-377 $c; done
for c in 16 17 18; do echo "== bls12-377 2^21 c=$c"; qt 21 bls12-377 $c; done
for c in 18 19 20; do echo "== bls12-377 2^23 c=$c"; qt 23 bls12-377 $c; done
for c in 17; do echo "== bls12-377 2^18 c=$c"; qt 18 bls12-377 $c; done
for c in 17 18; do echo "== pallas 2^18 c=$c"; qt 18 pallas $c; done
for c in 12 13 14; do echo "== pallas 2^16 c=$c"; qt 16 pallas $c; done
for c in 15 16 17 18; do echo "== pallas 2^20 c=$c"; qt 20 pallas $c; done
for c in 11 12 13 14; do echo "== ed-on-bls12-377 2^16 c=$c"; qt 16 ed-on-bls12-377 $c; done
for c in 15 16 17 18; do echo "== ed-on-bls12-377 2^20 c=$c"; qt 20 ed-on-bls12-377 $c; done
} > gpurun_out/r2j_sweeps.txt 2>&1
cat gpurun_out/r2j_sweeps.txt
