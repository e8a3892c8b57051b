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
  MGB_EMU_CXXFLAGS="-O1 -g -fsanitize=address -fno-omit-frame-pointer" \
  python -m pytest tests/test_host_emu_pipeline.py -q -p no:cacheprovider "$@"
rc=$?
ls /tmp/mgb_asan.log.* 2>/dev/null && head -40 /tmp/mgb_asan.log.*
exit $rc
