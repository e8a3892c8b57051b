#!/bin/bash
# Memcheck of the whole path WITHOUT a GPU: the emulated host build of tests/test_host_emu_pipeline.py (msm.cu against a
# stand-in CUDA runtime, every kernel on the SIMT emulator) compiled with AddressSanitizer.  "Device" buffers are heap
# blocks there, so every global-memory access of every kernel is checked against the size msm.cu allocated for it, and
# every host-side copy against the caller's buffers.  ~20 minutes on 8 cores; reports land in /tmp/mgb_asan.log.*.
# Round 2: 20 / 20 tests, no report from the product (the first run caught a test that passed a too short host buffer).
# The same with UndefinedBehaviorSanitizer (shifts by >= 32 bits and the like mean different things on the host and in
# PTX, so undefined behaviour in a kernel is a portability bug of the emulation AND a smell on the GPU): no report either --
#   MGB_EMU_CXXFLAGS="-O1 -g -fsanitize=undefined -fno-sanitize=alignment" python -m pytest tests/test_host_emu_pipeline.py -q
set -u
cd "$(dirname "$0")/.."
rm -f /tmp/mgb_asan.log.*
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:log_path=/tmp/mgb_asan.log \
  MGB_EMU_CXXFLAGS="-O1 -g -fsanitize=address -fno-omit-frame-pointer" MGB_TEST_FULL=1 \
  python -m pytest tests/test_host_emu_pipeline.py -q -p no:cacheprovider "$@"
rc=$?
# the production geometries (window 16: 2.6e5 buckets, 256 partial sums per group; tile sizes 56 / 28 / 14 / 64) on the same build
d=$(mktemp -d)
python tests/host_emu/make_emu_host.py montgomery_b200/csrc/msm.cu $d/msm_emu.cpp > /dev/null
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -shared -pthread -Wno-unknown-pragmas -DMGB_HOST_EMU -I tests/host_emu \
    -I montgomery_b200/csrc -I include -include cuda_rt_emu.h $d/msm_emu.cpp -o $d/libmgb_emu_asan.so -ldl
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:log_path=/tmp/mgb_asan.log python tests/host_emu/geometry_checks.py $d/libmgb_emu_asan.so || rc=1
ls /tmp/mgb_asan.log.* 2>/dev/null && head -40 /tmp/mgb_asan.log.*
exit $rc
