#!/usr/bin/env python
"""Benchmark of the proxmin NMF hot path on B200 (BASELINE.json metric and config).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A "step" is one PGM iteration of ``nmf.nmf`` on config 2 of BASELINE.json: Y = 8192 x 65536, K = 64,
prox_A = prox_plus, prox_S = prox_unity_plus (algorithms.py:87-135, nmf.py:28-65).  With N GPUs the
columns of Y and S are split over the ranks (strong scaling: the problem is fixed), A is replicated
and the G_A partials are summed with one NCCL all-reduce per iteration.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for how every number is obtained.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from proxmin_b200 import workloads  # noqa: E402

M_FULL, N_FULL, K_FULL = 8192, 65536, 64
METRIC = "NMF iterations/sec"
UNIT = "it/s"


def workload_name(M, N, K):
    return "nmf.nmf PGM Y=%dx%d K=%d prox_plus+prox_unity_plus" % (M, N, K)


def dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def measured_traffic(M, n_loc, K, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_grad_umma launch from the committed ncu capture
    (profiles/grad_umma_traffic.json); None when the capture is for another shape."""
    p = os.path.join(ROOT, "profiles", "grad_umma_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    w = t.get("workload", {})
    if (w.get("M"), w.get("N"), w.get("K")) != (M, n_loc, K):
        return None
    return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])


def init_gloo(world, rank):
    if world == 1:
        return None
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    return dist


def stripe_data(M, N, K, world, rank, seed=1234):
    """cfg2 recipe.  N>1: every rank draws the same A*, A0 and its own column stripe."""
    if world == 1:
        return workloads.cfg2(M, N, K, seed=seed) + ((0, N),)
    lo, hi = workloads.shard_columns(N, world, rank, align=128)
    n = hi - lo
    rng_a = np.random.default_rng(seed)
    At = rng_a.random((M, K), dtype=np.float32)
    A0 = rng_a.random((M, K), dtype=np.float32)
    rng = np.random.default_rng(seed + 1 + rank)
    St = rng.random((K, n), dtype=np.float32)
    Y = At @ St
    sd = np.float32(0.01 * 16.0)  # std(Y) of the recipe is ~16 for K=64 uniform factors; exact value is irrelevant here
    blk = max(1, (1 << 24) // max(n, 1))
    for r0 in range(0, M, blk):
        r1 = min(M, r0 + blk)
        Y[r0:r1] += sd * rng.standard_normal((r1 - r0, n), dtype=np.float32)
    np.maximum(Y, 0, out=Y)
    S0 = rng.random((K, n), dtype=np.float32)
    return Y, A0, S0, (lo, hi)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                return
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except Exception:
                continue
            for nme, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(M, N, K):
    # read Y once, read+write S, read+write A, fp32
    return 4 * (M * N + 2 * K * N + 2 * M * K)


def oracle_iterations_per_s(Y, A0, S0, iters, warm=1):
    """The reference algorithm (oracle port, same NumPy calls as nmf.py / algorithms.py) on the host cores."""
    from oracle import proxmin_oracle as orc

    stamps = []

    def cb(*X, it=None):
        stamps.append(time.perf_counter())

    A, S = A0.copy(), S0.copy()
    orc.nmf(Y, A, S, prox_A=orc.prox_plus, prox_S=orc.prox_unity_plus, algorithm="pgm", max_iter=warm + iters + 1,
            e_rel=0, callback=cb)
    # callback k fires at the start of iteration k: differences = whole iterations
    dt = np.diff(stamps)[warm:]
    return len(dt) / float(np.sum(dt)), float(np.median(dt))


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, world, rank):
    """--impl reference: the reference's CPU algorithm on a bounded column sample of the same workload."""
    if rank != 0:
        return
    M, N, K = args.M, args.N, args.K
    total = args.steps + args.warmup + 1
    # ~3.1 s per full-size fp32 iteration on 8 cores; keep the whole run under ~2 minutes
    est_full = 3.1 * (M * N * K) / (M_FULL * N_FULL * K_FULL)
    frac = min(1.0, 100.0 / max(total * est_full, 1e-9))
    Ns = max(128, int(N * frac) // 128 * 128)
    Ns = min(Ns, N)
    Y, A0, S0 = workloads.cfg2(M, Ns, K, seed=1234)
    ips, med = oracle_iterations_per_s(Y, A0, S0, args.steps, warm=args.warmup)
    value = ips * (Ns / N)  # cost per iteration is linear in the number of columns (O(MNK) + O(MN))
    cores = blas_threads()
    sample = "full M=%d, K=%d, first %d of %d columns per step; it/s scaled by %d/%d (cost linear in N)" % (
        M, K, Ns, N, Ns, N)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(M, N, K), "M": M, "N": N, "K": K},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def pinned_array(shape, dtype=np.float32):
    from proxmin_b200 import _ffi

    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    _ffi.check(_ffi.lib().pmx_host_alloc(n, ctypes.byref(p)))
    buf = (ctypes.c_char * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p


def run_product(args, world, rank, local):
    import proxmin_b200 as pmx
    from proxmin_b200 import _ffi
    from proxmin_b200 import nmf as pnmf

    M, N, K = args.M, args.N, args.K
    dist = init_gloo(world, rank)
    os.environ["PROXMIN_B200_DEVICE"] = str(local)
    ctx = _ffi.context()
    if world > 1:
        box = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        pmx.init_distributed(box[0], world, rank)

    Yg, A0, S0, (lo, hi) = stripe_data(M, N, K, world, rank)
    n_loc = hi - lo
    # host buffers of the public-API (e2e) leg live in pinned memory
    Y, y_ptr = pinned_array((M, n_loc))
    Y[...] = Yg
    del Yg

    chain_A = [(_ffi.OP_PLUS, 0, 0, 0.0)]
    chain_S = [(_ffi.OP_PLUS, 0, 0, 0.0), (_ffi.OP_UNITY, 0, 0, 0.0)]

    # ---------------- device-resident leg: `value` ----------------
    prob = pnmf.Problem(Y, A0, S0)
    prob.pgm_begin(chain_A, chain_S, accelerated=False, e_rel=(0.0, 0.0), kernel=args.kernel,
                   check_every=1 << 30)
    prob.pgm_run(args.warmup)
    ctx.sync()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(ctx.device)
    sampler.start()
    time.sleep(0.05)
    l0 = ctx.launches()
    ctx.sync()
    t0 = time.perf_counter()
    ctx.timer_start()
    done, _, _ = prob.pgm_run(args.steps)
    ms = ctx.timer_stop()
    ctx.sync()
    wall = time.perf_counter() - t0
    launches = ctx.launches() - l0
    assert done == args.steps, "timed region executed %d of %d iterations" % (done, args.steps)
    # second pass of the same length with per-launch CUDA events around the dominant kernel (on its own stream);
    # the steady-state CUDA-graph replay is off while events are recorded, so this pass is slightly slower
    ctx.profile(True)
    ctx.timer_start()
    done2, _, _ = prob.pgm_run(args.steps)
    ms_prof = ctx.timer_stop()
    kern_ms, kern_n = ctx.profile_read()
    ctx.profile(False)
    clocks = sampler.summary()
    if dist is not None:
        import torch

        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        dist.barrier()
    loss = prob.loss()
    prob.close()
    value = args.steps / (ms / 1e3)

    # ---------------- end-to-end leg through the public API with host buffers ----------------
    A_h, S_h = A0.copy(), S0.copy()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    pnmf.nmf(Y, A_h, S_h, prox_A=pmx.prox_plus, prox_S=pmx.prox_unity_plus, max_iter=args.steps, e_rel=0)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        import torch

        t = torch.tensor([e2e_s], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    h2d = (Y.nbytes + A0.nbytes + S0.nbytes) / args.steps
    d2h = (2 * A0.nbytes + 2 * S0.nbytes) / args.steps  # factors + last gradients handed back (algorithms.py:144)
    e2e = {"value": args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "note": "one nmf.nmf() solve of `steps` iterations from pinned host arrays: upload of Y,A,S + loop + "
                   "download of A,S,G inside the timed region"}

    if rank != 0:
        return
    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = measured_peaks()
    alg = algorithmic_bytes(M, n_loc, K)
    roof = None
    if kern_n > 0:
        avg_ms = kern_ms / kern_n
        ach = alg / (avg_ms * 1e-3) / 1e9
        flops = 6.0 * M * n_loc * K
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": measured_traffic(M, n_loc, K, world), "kernel": "k_grad_umma", "avg_launch_ms": avg_ms, "launches_timed": kern_n,
                "kernel_share_of_step": kern_ms / ms_prof, "peak_source": peak_src,
                "timing": "per-launch CUDA events in a second pass of `steps` iterations (graph replay off)",
                "algorithmic_bytes_per_launch": alg, "useful_tflops": flops / (avg_ms * 1e-3) / 1e12,
                "tensor_tflops_issued_bf16": 3 * flops / (avg_ms * 1e-3) / 1e12}
    # ---------------- CPU baseline (oracle port) on a bounded sample ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        Ns = min(N, max(128, (N // 16) // 128 * 128))
        Ys, As, Ss = workloads.cfg2(M, Ns, K, seed=1234)
        ips, med = oracle_iterations_per_s(Ys, As, Ss, 4, warm=1)
        cpu = {"value": ips * Ns / N, "unit": UNIT, "cores": blas_threads(), "kind": "port",
               "sample": "4 iterations on the first %d of %d columns (full M, K); it/s scaled by %d/%d" % (Ns, N, Ns, N)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(M, N, K), "M": M, "N": N, "K": K, "columns_per_gpu": n_loc,
                       "gemm": "tcgen05 kind::f16, 3-term bf16 split, fp32 accumulate" if args.kernel != 1 else "simt fp32",
                       "l2": "inputs larger than L2: every iteration streams the %.2f GB Y stripe" % (M * n_loc * 4 / 1e9)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "host_wall_ms_per_step": wall * 1e3 / args.steps, "final_loss": loss}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--M", type=int, default=M_FULL)
    ap.add_argument("--N", type=int, default=N_FULL)
    ap.add_argument("--K", type=int, default=K_FULL)
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 SIMT, 2 tcgen05")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    world, rank, local = dist_env()
    if args.gpus != world and world == 1 and args.gpus > 1:
        sys.stderr.write("bench.py: --gpus %d needs torchrun (one process per GPU); running 1 GPU\n" % args.gpus)
    if args.impl == "reference":
        run_reference(args, world, rank)
    else:
        run_product(args, world, rank, local)


if __name__ == "__main__":
    main()
