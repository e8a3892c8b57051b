#!/usr/bin/env python3
"""Benchmark of the MSM hot path (BASELINE.json: BLS12-377 G1, N = 2^20 points per GPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one complete MSM over one batch of synthetic input: fresh seeded scalars (a different
pre-generated set every step, as scripts/msm-weierstrass.ts:28-35 draws fresh scalars per run) against
a fixed seeded point set (known-dlog points a_i*G, the construction class of randomPointsFast,
src/curve-random.ts:24-92), result returned to the host as the canonical affine point.

  value   points/s with the scalars already resident in HBM (mgb_msm_device), whole job over all ranks
  e2e     the same through mgb_msm with HOST scalars in pinned memory: H2D of the scalars and D2H of
          the result inside the timed region.  The points stay resident, as they do in the
          reference (its `pointPtr` lives in wasm memory across calls; conversion is excluded from
          its timing, doc/zprize23.md:74).
  N > 1   one process per GPU (torchrun); points/scalars sharded contiguously, 2^20 pairs per GPU
          (weak scaling); each rank computes a partial sum, the partial accumulators are all-gathered
          with NCCL and every rank adds them and normalises.

The oracle (oracle/) is used here only for `cpu_baseline` and `--impl reference`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOGN_DEFAULT = 20
SEED_POINTS = 0x6D6F6E74
SEED_SCALARS = 0x6D6F6E74 ^ 3          # "seed xor config index" (config 4 = index 3)
ALGO_FIELD_MULTS_PER_POINT = 108.6     # SURVEY 8d, BLS12-377 2^20 at the reference's c = 18
MADS_PER_FIELD_MULT = 288              # 2 * 12^2 32x32->64 multiply-accumulates


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--logn", type=int, default=LOGN_DEFAULT, help="log2 of the points per GPU")
    ap.add_argument("--curve", default="bls12-377")
    ap.add_argument("--c", type=int, default=0, help="window bits (0 = engine default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-logn", type=int, default=0,
                    help="log2 size of the CPU baseline sample (0 = the whole workload when the host has >= 8 cores, else 2^18)")
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _measured_peak(key):
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh).get(key)
    except Exception:
        return None


def cpu_sample_logn(args):
    """The CPU arm runs the WHOLE per-GPU workload (one 2^20 MSM is ~2 s on 16 cores, ~30 core-seconds) unless the
    host is too small for that to stay within the bench's time budget."""
    if args.cpu_logn:
        return min(args.logn, args.cpu_logn)
    return args.logn if (os.cpu_count() or 1) >= 8 else min(args.logn, 18)


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by scripts/summarize_ncu.py traffic); None when absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


def dist_setup(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_reference_msm(label, points_bytes, scalars, n, threads):
    from oracle import cpu_ref
    res, ms = cpu_ref.msm(label, scalars, points_bytes, n, threads=threads)
    return res, ms


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle/msm_cpu.cpp, a native restatement --
    the reference's own Wasm needs Node, which this image does not have) on all host cores."""
    rank, world, local = dist_setup(args)
    if rank != 0:
        return
    from montgomery_b200 import curves, inputs
    curve = curves.BY_LABEL[args.curve]
    cl = cpu_sample_logn(args)
    n = 1 << cl
    # same seeded known-dlog point set as the GPU arm, generated on the CPU (untimed, like the
    # reference's randomPointsFast before its timed loop); no GPU code runs in this arm
    from oracle import cpu_ref
    threads = os.cpu_count() or 1
    pts = cpu_ref.known_dlog_points(args.curve, SEED_POINTS, n, threads)
    sets = [inputs.random_scalars(curve.q, n, SEED_SCALARS + i) for i in range(min(4, args.steps + args.warmup))]
    times = []
    for i in range(args.warmup + args.steps):
        _, ms = cpu_reference_msm(args.curve, pts, sets[i % len(sets)], n, threads)
        if i >= args.warmup:
            times.append(ms)
    ms_step = float(np.mean(times))
    value = n / (ms_step * 1e-3)
    sample = ("one MSM of 2^%d points per step (the whole per-GPU workload), %d threads" % (cl, threads)) if cl == args.logn else (
        "one MSM of 2^%d points per step (the first 2^%d of the 2^%d-per-GPU workload), %d threads" % (cl, cl, args.logn, threads))
    line = {
        "impl": "reference", "metric": "msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (native restatement; the reference computes in 29-bit limbs in i64)", "data": "synthetic",
        "config": {"workload": "%s MSM, 2^%d points per GPU, fresh scalars per step" % (args.curve, args.logn)},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_b200(args):
    import torch
    import montgomery_b200 as m
    from montgomery_b200 import _native, inputs
    import ctypes

    rank, world, local = dist_setup(args)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node = --gpus"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # stdout carries exactly one JSON line: NCCL prints its version banner with a plain printf when the communicator
        # is created, so file descriptor 1 points at stderr while that happens (init + one warm-up collective)
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    curve = m.curves.BY_LABEL[args.curve]
    n = 1 << args.logn
    from montgomery_b200.distributed import ShardedMsm
    sharded = ShardedMsm(curve, local, n)
    eng = sharded.engine
    sharded.random_points(n, SEED_POINTS)               # rank r: seed + r -> this rank's shard of the global point set
    nsets = 4
    host_sets = [torch.from_numpy(inputs.random_scalars(curve.q, n, SEED_SCALARS + 1000 * rank + i)).pin_memory() for i in range(nsets)]
    dev_sets = [h.cuda(non_blocking=False) for h in host_sets]
    torch.cuda.synchronize()
    opts_c = args.c or None

    def step(i, e2e):
        """one MSM over this rank's shard (+ all-gather and combine when world > 1)"""
        k = i % nsets
        return sharded.msm(host_sets[k] if e2e else dev_sets[k], n, on_device=not e2e, c=opts_c)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, sampler=None):
        for i in range(args.warmup):
            step(i, e2e)
        barrier()
        if sampler:
            sampler.start()
        t0 = time.perf_counter()
        phases = {}
        launches = 0
        res = None
        for i in range(args.steps):
            res, tm = step(args.warmup + i, e2e)
            launches += tm["n_launches"]
            for key in ("h2d_scalars", "decompose_slice", "sort", "accumulate", "reduce", "final_sum", "total"):
                phases[key] = phases.get(key, 0.0) + tm[key] / args.steps
        barrier()
        el = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([el], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), phases, launches, res, tm, clocks

    sampler = ClockSampler(local) if rank == 0 else None
    el, phases, launches, res, tm, clocks = timed(False, sampler)
    el_e2e, phases_e2e, _, res_e2e, _, _ = timed(True)
    ms_step = el / args.steps * 1e3
    total_points = n * world
    value = total_points / (el / args.steps)
    e2e_value = total_points / (el_e2e / args.steps)

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the path: integer-multiply pipe.  Peak measured live.
    lib = _native.lib()
    ops = ctypes.c_double()
    msb = ctypes.c_float()
    lib.mgb_microbench(local, 3, 2, 1024, 2000, ctypes.byref(ops), ctypes.byref(msb))    # carry-chained IMAD.WIDE.U32.X = one full 32x32+64->64 MAD
    peak_mads = ops.value
    lib.mgb_microbench(local, 0, 2, 1024, 2000, ctypes.byref(ops), ctypes.byref(msb))    # plain IMAD (mad.lo.u32) issue rate
    peak_imad_lo = ops.value
    algo_mads = ALGO_FIELD_MULTS_PER_POINT * MADS_PER_FIELD_MULT * n        # per GPU per step (reference op counts)
    acc_ms = phases["accumulate"]
    traffic = ncu_traffic() if (args.logn == LOGN_DEFAULT and args.curve == "bls12-377") else None
    achieved = algo_mads / (phases["total"] * 1e-3)
    roofline = {
        "bound": "imad", "achieved": achieved / 1e12, "peak": peak_mads / 1e12, "unit": "T 32x32->64 MAD/s",
        "frac": achieved / peak_mads, "traffic": traffic and traffic.get("bytes_per_launch"),
        "traffic_note": traffic and traffic.get("note"),
        "note": "integer-multiply roofline of BASELINE.md: algorithmic MADs (108.6 field mults/point x 288) / device time of the whole MSM "
                "(CUDA events on the engine's stream) / measured rate of carry-chained IMAD.WIDE.U32.X (one full 32x32+64->64 MAD per lane) on this GPU; "
                "plain IMAD (mad.lo) issues at %.2f T/s, i.e. a full MAD costs two IMAD slots, as SURVEY 8d assumed" % (peak_imad_lo / 1e12),
        "dominant_kernel": {"name": "k_batch_add", "phase_ms": acc_ms, "share_of_step": acc_ms / phases["total"],
                            "pairs_per_step": int(tm["n_pairs"]), "field_mults_per_s": 6.0 * tm["n_pairs"] / (acc_ms * 1e-3)},
    }
    # HBM side of the path (north star: "HBM GB/s for the sort and gather phases").  The sort no longer copies points:
    # per sorted entry the scatter reads digit + rank (8 B) and the bucket's offset + count (8 B, L2-resident) and writes
    # a 4-byte point reference (+ half a byte of tree depth); the 96-byte gather of every point happens inside round 0 of
    # k_batch_add, straight from the point table (2 x 96 B read + 96 B written per addition of round 0).
    ent_bytes = 8 + 8 + 4 + 0.5
    hbm = {"sort_scatter_gbs": tm["n_pairs"] and (n * 2 * tm["K"] * ent_bytes) / (phases["sort"] * 1e-3) / 1e9,
           "sort_ms": phases["sort"], "hbm_peak_gbs": _measured_peak("hbm_gbs"),
           "note": "sort = 3 scan kernels + scatter of 4-byte references; latency/atomic-bound, not bandwidth-bound at this size"}
    line = {
        "metric": "msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (12 x 32-bit Montgomery)", "data": "synthetic",
        "config": {"workload": "%s G1 MSM, 2^%d points per GPU, fresh scalars per step, points resident" % (args.curve, args.logn),
                   "window_bits": tm["c"], "windows": tm["K"], "l2": "inputs larger than L2 (point table %d MB + scalars %d MB per GPU)" % (
                       n * 144 >> 20, n * 32 >> 20), "parallelism": "points sharded x%d, NCCL all-gather of partial sums" % world},
        "msm_ms": ms_step, "phases_ms": phases, "result_x": hex(res["x"]),
        "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": el_e2e / args.steps * 1e3,
                "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": (2 * curve.coord_bytes + 4) * world, "phases_ms": phases_e2e},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "hbm_phases": hbm,
    }
    if world == 1 and not args.no_cpu_baseline:
        cl = cpu_sample_logn(args)
        ncpu = 1 << cl
        pts, _ = eng.get_points(0, ncpu)
        threads = os.cpu_count() or 1
        sc = host_sets[0].numpy()[:ncpu]
        cres, cms = cpu_reference_msm(args.curve, pts, sc, ncpu, threads)
        gres, _ = eng.msm(sc, n=ncpu)
        line["cpu_baseline"] = {"value": ncpu / (cms * 1e-3), "unit": "points/s", "cores": threads, "kind": "port",
                                "sample": ("one MSM of the whole 2^%d-point workload, %.0f ms" if cl == args.logn else
                                           "one MSM of the first 2^%d points of the workload, %.0f ms") % (cl, cms),
                                "agrees_with_gpu": cres == gres}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
