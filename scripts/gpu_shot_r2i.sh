#!/bin/bash
# round 2, GPU call 9: scatter with one (offset, count) gather; window / depth sweeps at the small BASELINE configs
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.txt 2>&1; tail -3 gpurun_out/r2i_pytest.txt
qt() { timeout 60 python scripts/quick_time.py "$@" 2>&1 | tail -1 | sed -E "s/.*('decompose_slice': [0-9.]+).*('sort': [0-9.]+).*('accumulate': [0-9.]+).*('reduce': [0-9.]+).*('final_sum': [0-9.]+).*('total': [0-9.]+).*('c': [0-9]+).*('K': [0-9]+).*('rounds': [0-9]+).*/\1 \2 \3 \4 \5 \6 \7 \8 \9/"; }
{
echo "== 2^20 default"; qt 20; qt 20
for nr in 6 7; do echo "== 2^20 NROUNDS=$nr"; MGB_DEBUG_NROUNDS=$nr qt 20; done
for c in 10 11 12 13 14 15; do echo "== bls12-377 2^16 c=$c"; qt 16 bls12-377 $c; done
for c in 12 13 14 15 16; do echo "== pallas 2^18 c=$c"; qt 18 pallas $c; done
for c in 12 13 14 15 16; do echo "== ed-on-bls12-377 2^18 c=$c"; qt 18 ed-on-bls12-377 $c; done
for c in 12 13 14 15 16; do echo "== bls12-377 2^18 c=$c"; qt 18 bls12-377 $c; done
for c in 17 18 19 20; do echo "== bls12-377 2^22 c=$c"; qt 22 bls12-377 $c; done
for c in 10 11 12; do echo "== bls12-377 2^14 c=$c"; qt 14 bls12-377 $c; done
} > gpurun_out/r2i_sweeps.txt 2>&1
cat gpurun_out/r2i_sweeps.txt
