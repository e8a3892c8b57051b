"""Manual tool (not collected by pytest; ~6 minutes): runs bench.py's CUDA arm end to end WITHOUT a GPU, on the emulated
host build of the engine (tests/host_emu: msm.cu against a stand-in CUDA runtime, every kernel on the SIMT emulator), to
check the bench script itself -- every block of the JSON line (headline, `configs` with the CPU port beside them,
`strong_2p24`, `cpu_baseline`, roofline arithmetic, parity bookkeeping) -- after an edit, before GPU time is spent on it.
Sizes are shrunk (2^7 / 2^6 / 2^10 points), torch's CUDA calls are stubbed ("device" memory is host memory in the
emulation), the microbenchmarks return a constant.  Timings mean nothing; `parity_ok` and the structure do.

    python tests/bench_dry_run.py            # prints the JSON line; exit code 0 = the whole script ran and parity held
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EMU = os.path.join(ROOT, "tests", "host_emu")

STUB = r"""
#include <cstddef>
#include <cstdint>
extern "C" int mgb_field_op(int, int, int, const uint8_t*, const uint8_t*, uint8_t*, size_t) { return -1; }
extern "C" int mgb_microbench(int, int, int, int, int, double* ops, float* ms) { *ops = 1e12; *ms = 1.0f; return 0; }
"""


def build(d):
    src, stub, so = os.path.join(d, "msm_emu.cpp"), os.path.join(d, "stub.cpp"), os.path.join(d, "libmgb_emu.so")
    subprocess.check_call([sys.executable, os.path.join(EMU, "make_emu_host.py"), os.path.join(ROOT, "montgomery_b200", "csrc", "msm.cu"), src])
    open(stub, "w").write(STUB)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-DMGB_HOST_EMU", "-I", EMU,
                           "-I", os.path.join(ROOT, "montgomery_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           "-include", "cuda_rt_emu.h", src, stub, "-o", so, "-ldl"])
    return so


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as d:
        os.environ["MGB_LIB"] = build(d)          # read by montgomery_b200/_native.py at import
        import torch
        torch.cuda.set_device = lambda *a, **k: None
        torch.cuda.synchronize = lambda *a, **k: None
        torch.Tensor.cuda = lambda self, *a, **k: self.clone()
        torch.Tensor.pin_memory = lambda self, *a, **k: self
        _tensor = torch.tensor
        torch.tensor = lambda *a, device=None, **k: _tensor(*a, **k)
        import bench
        bench.LOGN_DEFAULT = 7
        bench.STRONG_LOGN = 10                    # the strong block runs when 2^STRONG_LOGN / world >= 2^10
        for cfg in bench.EXTRA_CONFIGS.values():
            cfg["logn"] = 6
        sys.argv = ["bench.py", "--steps", "1", "--warmup", "0", "--logn", "7"]
        bench.main()
