"""Manual tool (not collected by pytest; ~6 minutes): runs bench.py's CUDA arm end to end WITHOUT a GPU, on the emulated
host build of the engine (tests/host_emu: msm.cu against a stand-in CUDA runtime, every kernel on the SIMT emulator), to
check the bench script itself -- every block of the JSON line (headline, `configs` with the CPU port beside them,
`strong_2p24`, `cpu_baseline`, roofline arithmetic, parity bookkeeping) -- after an edit, before GPU time is spent on it.
Sizes are shrunk (2^7 / 2^6 / 2^10 points), torch's CUDA calls are stubbed ("device" memory is host memory in the
emulation), the microbenchmarks return a constant.  Timings mean nothing; `parity_ok` and the structure do.

    python tests/bench_dry_run.py            # prints the JSON line; exit code 0 = the whole script ran and parity held
    python tests/bench_dry_run.py --ranks 2  # the multi-GPU shape: one process per rank as torchrun starts them,
                                             # torch.distributed on gloo, the engine's own collective over the
                                             # file rendezvous of tests/host_emu/fake_nccl.cpp (~15 minutes)
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
    subprocess.check_call([sys.executable, os.path.join(EMU, "make_emu_host.py"), os.path.join(ROOT, "montgomery_b200", "csrc", "msm.cu"), src], stdout=sys.stderr)
    open(stub, "w").write(STUB)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-DMGB_HOST_EMU", "-I", EMU,
                           "-I", os.path.join(ROOT, "montgomery_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           "-include", "cuda_rt_emu.h", src, stub, "-o", so, "-ldl"])
    return so


def run_ranks(world):
    """parent of a multi-rank dry run: builds once, starts one child per rank, relays rank 0's JSON line"""
    with tempfile.TemporaryDirectory() as d:
        so = build(d)
        nccl = os.path.join(d, "libfake_nccl.so")
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", os.path.join(EMU, "fake_nccl.cpp"), "-o", nccl])
        os.makedirs(os.path.join(d, "rendezvous"))
        procs = []
        for r in range(world):
            env = dict(os.environ, MGB_LIB=so, MGB_NCCL_LIB=nccl, MGB_FAKE_NCCL_DIR=os.path.join(d, "rendezvous"), RANK=str(r), LOCAL_RANK=str(r),
                       WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29400 + os.getpid() % 500), MGB_DRY_RUN_CHILD="1")
            procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__)], env=env, stdout=subprocess.PIPE, text=True))
        outs = [p.communicate()[0] for p in procs]
        print(outs[0].strip())
        return max(p.returncode for p in procs)


if __name__ == "__main__":
    if "--ranks" in sys.argv:
        sys.exit(run_ranks(int(sys.argv[sys.argv.index("--ranks") + 1])))
    with tempfile.TemporaryDirectory() as d:
        child = bool(os.environ.get("MGB_DRY_RUN_CHILD"))
        if not child:
            os.environ["MGB_LIB"] = build(d)      # read by montgomery_b200/_native.py at import
        import torch
        torch.cuda.set_device = lambda *a, **k: None
        torch.cuda.synchronize = lambda *a, **k: None
        torch.Tensor.cuda = lambda self, *a, **k: self.clone()
        torch.Tensor.pin_memory = lambda self, *a, **k: self
        _tensor = torch.tensor
        torch.tensor = lambda *a, device=None, **k: _tensor(*a, **k)
        world = int(os.environ.get("WORLD_SIZE", "1")) if child else 1
        if world > 1:
            import torch.distributed as dist
            _init = dist.init_process_group
            dist.init_process_group = lambda backend=None, **k: _init("gloo")      # no device_id: CPU tensors
        import bench
        bench.LOGN_DEFAULT = 7
        bench.STRONG_LOGN = 10 + (world.bit_length() - 1)      # the strong block runs when 2^STRONG_LOGN / world >= 2^10
        for cfg in bench.EXTRA_CONFIGS.values():
            cfg["logn"] = 6
        sys.argv = ["bench.py", "--gpus", str(world), "--steps", "1", "--warmup", "0", "--logn", "7"]
        bench.main()
