#!/usr/bin/env python
"""bench.py -- spread+predict throughput of the B200 hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU algorithm

Workload (config.workload = "C4"): BASELINE.json configs[3] / SURVEY.md 8(d) -- the configuration
the target metric is quoted on and which fits one B200: 100 000 queries x 50 000 targets, 20 000
sources = 20 000 similarity features, weighted features at alpha = 0 (100 % dense), labels
Bernoulli(0.05), FP64.  One step = degrees -> spread -> T = (Xs' * (Y ./ ks)) ./ kf -> R = Xq * T
with clean! fused, from featurized operands resident in HBM to R resident in HBM.
One score = one entry of the Nq x Nt result.  At N > 1 the same problem is sharded (strong scaling):
query rows of Xq / R by rank, T by target-column block with one NCCL all-gather.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "predicted query-target scores/sec (spread+predict)"
UNIT = "scores/s"
C4 = dict(nq=100_000, ns=20_000, nf=20_000, nt=50_000, y_density=0.05, alpha=0.0, weighted=True)
SEED = 20244
NOMINAL_FP64_TFLOPS = 37.0  # HGX B200 datasheet: 296 TFLOP/s FP64 tensor per 8 GPUs


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every dimension (debug only; "
                    "a scaled run is not a bench value)")
    ap.add_argument("--precision", choices=["f64", "tf32", "f64_int8"], default="f64",
                    help="opt-in reduced-precision chain on tcgen05 (not the BASELINE metric; N=1 only)")
    ap.add_argument("--reference-workload", choices=["C4", "C2", "C3"], default="C4",
                    help="with --impl reference: which BASELINE config the CPU restatement is timed on (C4 = the bench "
                         "workload; C2 / C3 are the side measurements quoted in DESIGN.md)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip e2e_api and the side configs (C2 / C3 / C5)")
    return ap.parse_args()


def dims(scale: float):
    d = dict(C4)
    if scale != 1.0:
        for k_ in ("nq", "ns", "nf", "nt"):
            d[k_] = max(16 * 8, int(round(d[k_] * scale / 128)) * 128)
    return d


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's literal restatement of the reference's CPU path
# ------------------------------------------------------------------------------------------------


def host_threads() -> int:
    """Host cores this process may use (the affinity mask, not os.cpu_count())."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def force_blas_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to its workers, which would time the CPU arm on one core.  The CPU arm
    uses every host core it may run on: the environment is overridden (in case NumPy is not loaded yet) and the
    BLAS pool is resized explicitly (in case it is).  Returns the thread count actually in use."""
    n = host_threads()
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(n)
    import numpy  # noqa: F401  (loads the BLAS whose pool is resized below)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=n)
        blas = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return int(max(blas)) if blas else n
    except Exception:
        return n


def cpu_dgemm_rate(np, m: int = 2048) -> float:
    """flop/s of the host BLAS on an m^3 DGEMM (used only to size the bounded sample)."""
    a = np.random.default_rng(1).random((m, m))
    best = 0.0
    for _ in range(6):  # the BLAS pool needs a few calls to spin up after a resize
        t0 = time.perf_counter()
        a @ a
        best = max(best, 2.0 * m ** 3 / (time.perf_counter() - t0))
    return best


def cpu_block_reduced_full_inner(np, threads: int):
    """Block-reduced CPU form (T = (Xs' * (Y ./ ks)) ./ kf, R = Xq * T) in NumPy/OpenBLAS at the FULL
    inner dimensions of C4 (Ns = Nf = 20 000), on a sample of query rows and target columns.  A score costs the same
    flops here as in the full problem once the T product is charged per target column and the R product per score:
    t_full = t_T * Nt / nt_s + t_R * (Nq * Nt) / (nq_s * nt_s).  A proxy for "a sane CPU implementation" -- the
    reference itself runs the literal n x n path (timed as `value`)."""
    ns = nf = C4["ns"]
    nq_s, nt_s = 10_000, 2_500
    rng = np.random.default_rng(SEED)
    Xs = rng.random((ns, nf))
    Xq = rng.random((nq_s, nf))
    Y = (rng.random((ns, nt_s)) < C4["y_density"]).astype(np.float64)
    t0 = time.perf_counter()
    ks = np.count_nonzero(Xs, axis=1) + np.count_nonzero(Y, axis=1)
    kf = np.count_nonzero(Xs, axis=0)
    Wst = Y / np.maximum(ks, 1)[:, None]
    T = (Xs.T @ Wst) / np.maximum(kf, 1)[:, None]
    t_T = time.perf_counter() - t0
    t0 = time.perf_counter()
    R = Xq @ T
    t_R = time.perf_counter() - t0
    del R
    t_full = t_T * C4["nt"] / nt_s + t_R * (C4["nq"] * C4["nt"]) / (nq_s * nt_s)
    return {"value": C4["nq"] * C4["nt"] / t_full, "unit": UNIT, "cores": threads,
            "shape": f"Ns = Nf = {ns} (full), sample of {nq_s} query rows x {nt_s} target columns",
            "t_T_s": t_T, "t_R_s": t_R, "extrapolated_full_step_s": t_full,
            "note": "block-reduced NumPy/OpenBLAS chain at C4's full inner dimensions (48 000 flop per score as on "
                    "the GPU); extrapolated linearly in target columns (T) and scores (R); NOT the reference's path"}


def cpu_reference_run(steps: int, warmup: int, budget_s: float = 100.0, with_block_reduced: bool = True):
    """Times `construct -> spread -> A*(W*W) -> slice -> clean!` (reference src/core.jl:148-201, 365-371, 402-423,
    478-484) in NumPy/OpenBLAS on every host core, on a 1/div-scale replica of C4 (the literal n x n path needs
    289 GB per matrix at full size and cannot run).  `div` is chosen from a DGEMM calibration so that the
    warmup + steps passes fit `budget_s`.  Returns (cpu_baseline dict, seconds per step)."""
    threads = force_blas_threads()
    import numpy as np
    from oracle import simspread_oracle as o
    rate = cpu_dgemm_rate(np)
    per_step = budget_s / max(1, steps + warmup)
    n_full = C4["nq"] + C4["ns"] + C4["nf"] + C4["nt"]
    n_fit = (0.6 * per_step * rate / 4.0) ** (1.0 / 3.0)       # literal path: two n^3 DGEMMs = 4 n^3 flop
    sample_div = int(min(80, max(16, -(-n_full // max(1.0, n_fit)))))
    nq, ns, nf, nt = (C4["nq"] // sample_div, C4["ns"] // sample_div, C4["nf"] // sample_div,
                      C4["nt"] // sample_div)
    Xq, Xs, Y = o.synth_dense(nq, ns, nf, nt, seed=SEED, y_density=C4["y_density"], alpha=C4["alpha"],
                              weighted=C4["weighted"])
    names = [str(i) for i in range(nq + ns + nf + nt)]
    rows, cols = names[:nq], names[nq + ns + nf:]

    def step():
        A = o._assemble4(Xq, Xs, Y)  # construct: dense 4-layer adjacency matrix
        B = A.copy()
        B[:nq, :] = 0.0
        B[:, :nq] = 0.0
        R = o.predict_dense(A, B, names, rows, cols)
        o.clean(R, A, names, cols)
        return R

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    n = nq + ns + nf + nt
    base = {
        "value": nq * nt / t, "unit": UNIT, "cores": int(threads), "kind": "port",
        "sample": f"1/{sample_div}-scale C4 replica (nq={nq}, ns=nf={ns}, nt={nt}; dense n={n}), literal "
                  f"NumPy/OpenBLAS restatement of construct+spread+A*(W*W)+clean!, median of {steps}, {threads} BLAS "
                  f"threads (host DGEMM {rate / 1e9:.0f} GFLOP/s); Julia is not installed, this is a port not "
                  "SimSpread.jl itself",
        "seconds_per_step": t,
    }
    if with_block_reduced:
        try:
            base["block_reduced"] = cpu_block_reduced_full_inner(np, threads)
            base["block_reduced_value"] = base["block_reduced"]["value"]
        except MemoryError:
            base["block_reduced"] = {"error": "not enough host memory for the full-inner-dimension sample"}
    return base, t


def synth_small_config(np, which: str):
    """BASELINE config 2 (Enzyme-shaped: 445 x 664, Beta(2,5) similarities, binary alpha = 0.35, 10 folds) or config 3
    (5 000 queries + 5 000 sources x 2 000 targets, uniform similarities, weighted); seeded."""
    from oracle import simspread_oracle as o
    rng = np.random.default_rng(20241)
    if which == "C2":
        N, Nt, alpha, weighted = 445, 664, 0.35, False
        S = np.round(rng.beta(2, 5, size=(N, N)), 6)
        np.fill_diagonal(S, 1.0)
        Y = (rng.random((N, Nt)) < 0.0099).astype(float)
        names = [f"D{i:04d}" for i in range(N)]
        folds = o.split_round_robin([names[i] for i in rng.permutation(N)], 10)
    else:
        nq, ns, Nt, alpha, weighted = 5000, 5000, 2000, 0.5, True
        N = nq + ns
        S = np.round(rng.random((N, N)), 6)
        Y = (rng.random((N, Nt)) < 0.01).astype(float)
        names = [f"n{i}" for i in range(N)]
        folds = [names[:nq]]
    tn = [f"t{j}" for j in range(Nt)]
    return S, Y, names, tn, folds, alpha, weighted


def cpu_reference_small_configs(which: str):
    """The reference's literal CPU path (dense n x n construct -> spread -> A*(W*W) -> clean!) on BASELINE config 2
    (Enzyme-shaped, all 10 folds) or on ONE alpha of config 3 (n = 17 000; the dense DGEMMs do not depend on alpha)."""
    threads = force_blas_threads()
    import numpy as np
    from oracle import simspread_oracle as o
    S, Y, names, tn, folds, alpha, weighted = synth_small_config(np, which)
    nq_total, Nt = sum(len(f) for f in folds), len(tn)
    t0 = time.perf_counter()
    Xo, xr, xc = o.featurize(S, names, names, alpha, weighted)
    n_full = 0
    for q in folds:
        Ao, Bo, nn = o.construct_queries(Y, (names, tn), Xo, (xr, xc), q)
        w = o.predict_dense(Ao, Bo, nn, q, tn)
        o.clean(w, Ao, nn, tn)
        n_full = Ao.shape[0]
    t = time.perf_counter() - t0
    return {"value": nq_total * Nt / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{which}: literal NumPy/OpenBLAS restatement of featurize + construct + spread + A*(W*W) + clean!, "
                      f"{len(folds)} fold(s), dense n = {n_full}", "seconds_per_step": t}, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.reference_workload != "C4":
        base, t = cpu_reference_small_configs(args.reference_workload)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
                          "steps": 1, "warmup": 0, "ms_per_step": t * 1e3, "higher_is_better": True, "dtype": "f64",
                          "data": "synthetic", "config": {"workload": args.reference_workload}, "cpu_baseline": base,
                          "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return 0
    base, t = cpu_reference_run(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4", **{k_: C4[k_] for k_ in ("nq", "ns", "nf", "nt")},
                   "note": "reference CPU algorithm on a bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id: str):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", gpu_id, f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def gpu_smi_id(torch, local: int) -> str:
    """nvidia-smi selector of the CUDA device `local` (UUID when torch exposes it)."""
    try:
        u = str(torch.cuda.get_device_properties(local).uuid)
        return u if u.startswith("GPU-") else "GPU-" + u
    except Exception:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if local < len(ids):
                return ids[local]
        return str(local)


def colmajor(torch, rows, cols, device):
    """torch buffer holding a column-major rows x cols float64 matrix with ld = round_up(rows, 16)."""
    ld = (rows + 15) // 16 * 16
    return torch.zeros((cols, ld), dtype=torch.float64, device=device), ld


def fill_uniform6(torch, buf, rows, seed):
    g = torch.Generator(device=buf.device)
    g.manual_seed(seed)
    step = max(1, (1 << 27) // buf.shape[1])
    for c0 in range(0, buf.shape[0], step):
        blk = buf[c0:c0 + step, :rows]
        blk.copy_(torch.round(torch.rand(blk.shape, generator=g, device=buf.device, dtype=torch.float64) * 1e6) / 1e6)


def fill_bernoulli(torch, buf, rows, p, seed):
    g = torch.Generator(device=buf.device)
    g.manual_seed(seed)
    step = max(1, (1 << 27) // buf.shape[1])
    for c0 in range(0, buf.shape[0], step):
        blk = buf[c0:c0 + step, :rows]
        blk.copy_((torch.rand(blk.shape, generator=g, device=buf.device) < p).to(torch.float64))


def fp64_gemm_peak(torch, device):
    """Measured cuBLAS DGEMM throughput on this GPU (MEASURED_PEAKS.json has no FP64 entry)."""
    n = 8192
    a = torch.rand((n, n), dtype=torch.float64, device=device)
    b = torch.rand((n, n), dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    torch.cuda.empty_cache()
    return best


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import simspread_b200 as ss
    from simspread_b200._lib import SS_OP_N, SS_OP_T, SS_PREDICT_CLEAN, check

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ss.build()
    ctx = ss.Context(local)  # raises without a B200 -- there is no CPU fallback
    ss.Context._default = ctx
    L = ss.lib()
    d = dims(args.scale)
    nq, ns, nf, nt = d["nq"], d["ns"], d["nf"], d["nt"]
    assert nq % world == 0 and nt % world == 0, "query rows / target columns must divide over the ranks"
    nq_l, nt_l = nq // world, nt // world
    q0, t0 = rank * nq_l, rank * nt_l

    # ---- synthetic operands, generated on the device and featurized by the library (untimed) ----
    bXq, ldq = colmajor(torch, nq_l, nf, dev)
    bXs, lds = colmajor(torch, ns, nf, dev)
    bY, ldy = colmajor(torch, ns, nt_l, dev)
    fill_uniform6(torch, bXq, nq_l, SEED + 1000 + rank)  # query similarities (row slab of this rank)
    fill_uniform6(torch, bXs, ns, SEED + 1)              # source similarities (replicated)
    fill_bernoulli(torch, bY, ns, d["y_density"], SEED + 2000 + rank)  # label block of this rank
    torch.cuda.synchronize()
    mXq = ss.DMat.wrap(ctx, bXq.data_ptr(), nq_l, nf, ldq)
    mXs = ss.DMat.wrap(ctx, bXs.data_ptr(), ns, nf, lds)
    mY = ss.DMat.wrap(ctx, bY.data_ptr(), ns, nt_l, ldy)
    check(L.ss_featurize(ctx.h, mXq.h, d["alpha"], int(d["weighted"]), mXq.h))
    check(L.ss_featurize(ctx.h, mXs.h, d["alpha"], int(d["weighted"]), mXs.h))
    bR, ldr = colmajor(torch, nq_l, nt, dev)
    mR = ss.DMat.wrap(ctx, bR.data_ptr(), nq_l, nt, ldr)

    # N > 1: every exchange step lives behind the C ABI (ss_comm_* / ss_predict_query_sharded: NCCL dlopen()ed by the
    # library, T tiles stored into the peers from the GEMM epilogue).  torch.distributed only launched the ranks and
    # carries the 128-byte NCCL unique id to them.
    sharded = comm = None
    if world > 1:
        from simspread_b200.sharded import Comm, ShardedQuery

        def exchange(raw: bytes) -> bytes:
            box = [raw]
            dist.broadcast_object_list(box, src=0)
            return box[0]

        comm = Comm(ctx, rank, world, exchange=exchange)
        sharded = ShardedQuery(comm, ns, nf, nt)
        assert sharded.nt_blk == nt_l

    ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    pflag = {"f64": 0, "tf32": 1 << 4, "f64_int8": 3 << 4}[args.precision]
    if pflag and world > 1:
        raise SystemExit("--precision tf32 is wired for --gpus 1 only")

    def step():
        if world == 1:
            check(L.ss_predict_query(ctx.h, mXq.h, mXs.h, mY.h, mR.h, SS_PREDICT_CLEAN | pflag, None))
        else:
            sharded.predict(mXq, mXs, mY, mR, clean=True)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            comm.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(gpu_smi_id(torch, local))
    ctx.profile(True)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        step()
    e1.record(ext)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launch_count() - launches0
    if world > 1:
        ms_total = comm.allreduce_host([ms_total], "max")[0]
    ms_step = ms_total / args.steps
    value = nq * nt / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (R = Xq * T) ---------------------------------------------
    r_flops = 2.0 * nq_l * nt * nf
    if pflag:  # opt-in modes: report the GEMM times only, no FP64 roofline claim
        args.no_e2e = True
    r_times = [ms for ms, fl in prof if abs(fl - r_flops) < 0.5]
    t_times = [ms for ms, fl in prof if abs(fl - r_flops) >= 0.5]
    # which form the first product took: the dense DMMA GEMM records 2 Ns Nf Nt flop, the edge-list form of a sparse label
    # matrix (csrc/ss_tsparse.cu) 2 nnz(Y) Nf -- every entry of T is computed either way
    t_fl = [fl for ms, fl in prof if abs(fl - r_flops) >= 0.5]
    dense_t_flops = 2.0 * ns * nf * (nt // world if world > 1 else nt)
    t_form = None
    if t_fl:
        t_form = ("dense DMMA GEMM (ss_dgemm_whole_kernel<!A_MMAJOR>)" if t_fl[0] > 0.5 * dense_t_flops else
                  "edge list of the label matrix, %.1f %% dense (tsp_kernel: CSC of W by target column, FP64 FMA; "
                  "SS_T_FORM=dense forces the DMMA GEMM)" % (100.0 * t_fl[0] / dense_t_flops))
    peak_meas = fp64_gemm_peak(torch, dev)
    roofline = None
    if r_times:
        achieved = r_flops / (statistics.mean(r_times) * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
        if os.path.exists(tpath) and args.scale == 1.0 and world == 1:
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {
            "bound": "tensor", "kernel": "ss_dgemm_whole_kernel<A_MMAJOR> (R = Xq*T, FP64 DMMA; whole 128 x 128 tiles)",
            "achieved": achieved, "peak": peak_meas, "unit": "TFLOP/s", "frac": achieved / peak_meas,
            "traffic": traffic,
            "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul float64) measured in this run, best of 5; "
                           "MEASURED_PEAKS.json has no FP64 entry",
            "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_nominal": achieved / NOMINAL_FP64_TFLOPS,
            "flops_per_launch": r_flops, "ms_per_launch": statistics.mean(r_times),
            "share_of_step": statistics.mean(r_times) / ms_step,
            "t_gemm_ms": statistics.mean(t_times) if t_times else None,
            "t_product": t_form,
        }

    # ---- parity spot check at full size (untimed): sampled entries recomputed with torch/cuBLAS -----
    checkres = None
    if not args.no_check and world == 1:
        g = torch.Generator(device="cpu")
        g.manual_seed(7)
        tq = torch.randint(0, nq_l, (64,), generator=g).to(dev)
        tt = torch.randint(0, nt, (64,), generator=g).to(dev)
        Xs_v = bXs[:, :ns]            # (nf, ns) == Xs'
        Y_v = bY[:, :ns]              # (nt, ns) == Y'
        ks_ = (Xs_v != 0).sum(0) + (Y_v != 0).sum(0)
        kf_ = (Xs_v != 0).sum(1)
        kt_ = (Y_v != 0).sum(1)
        Wst_cols = torch.nan_to_num(Y_v[tt] / ks_.to(torch.float64), nan=0.0, posinf=0.0)   # (64, ns)
        Tcols = (Xs_v @ Wst_cols.T)                                                           # (nf, 64)
        Tcols = torch.where(kf_[:, None] > 0, Tcols / kf_[:, None].to(torch.float64), torch.zeros_like(Tcols))
        want = (bXq[:, :nq_l].T[tq] * Tcols.T).sum(1)                                          # (64,)
        want = torch.where(kt_[tt] == 0, torch.full_like(want, -99.0), want)
        got = bR[tt, tq]
        rel = ((got - want).abs() / want.abs().clamp_min(1e-300)).max().item()
        tol = {"f64": 1e-12, "tf32": 2e-3, "f64_int8": 1e-12}[args.precision]
        checkres = {"sampled_entries": 64, "max_rel_err": rel, "tolerance": tol, "ok": bool(rel < tol)}

    if not args.no_check and world > 1:
        # (i) every rank must hold the same assembled T; (ii) entries of this rank's R slab whose target column lies
        # in this rank's Y block are recomputed with torch/cuBLAS (torch.distributed only sums the checker's degrees)
        from simspread_b200.sharded import _CudaView
        hT, _hk = sharded.views()
        tr_, tc_, tld_, tp_ = C.c_int64(), C.c_int64(), C.c_int64(), C.c_void_p()
        check(L.ss_mat_info(hT, C.byref(tr_), C.byref(tc_), C.byref(tld_), C.byref(tp_)))
        bT = torch.as_tensor(_CudaView(tp_.value, (tc_.value, tld_.value)), device=dev)[:, :nf]
        cs = comm.allreduce_host([float(bT.sum().item()), -float(bT.sum().item())], "max")
        same_T = bool(cs[0] == -cs[1])  # max(x) == min(x) over the ranks
        g = torch.Generator(device="cpu")
        g.manual_seed(11 + rank)
        tq = torch.randint(0, nq_l, (32,), generator=g).to(dev)
        tl = torch.randint(0, nt_l, (32,), generator=g).to(dev)
        Xs_v, Y_v = bXs[:, :ns], bY[:, :ns]
        ky = (Y_v != 0).sum(0).to(torch.int64)
        dist.all_reduce(ky)
        ks_ = (Xs_v != 0).sum(0) + ky
        kf_ = (Xs_v != 0).sum(1)
        kt_ = (Y_v != 0).sum(1)
        Wst_cols = torch.nan_to_num(Y_v[tl] / ks_.to(torch.float64), nan=0.0, posinf=0.0)
        Tcols = Xs_v @ Wst_cols.T
        Tcols = torch.where(kf_[:, None] > 0, Tcols / kf_[:, None].to(torch.float64), torch.zeros_like(Tcols))
        want = (bXq[:, :nq_l].T[tq] * Tcols.T).sum(1)
        want = torch.where(kt_[tl] == 0, torch.full_like(want, -99.0), want)
        got = bR[tl + t0, tq]
        rel = float(((got - want).abs() / want.abs().clamp_min(1e-300)).max().item())
        rel = comm.allreduce_host([rel], "max")[0]
        checkres = {"sampled_entries": 32 * world, "max_rel_err": rel, "tolerance": 1e-12,
                    "ok": bool(rel < 1e-12 and same_T), "all_ranks_hold_identical_T": same_T}

    # ---- e2e: reference-facing host-buffer call, H2D/D2H inside the timed region ----------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, ss, ctx, torch, comm, np, dev, world, rank, d, nq_l, nt_l, bXq, bXs, bY, ldq, lds, ldy,
                      sharded,
                      (mXq, mXs, mY, mR, bR, ldr))

    # ---- the line (everything the contract asks for is known at this point) --------------------------------
    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu, _ = cpu_reference_run(steps=3, warmup=1, budget_s=30.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "C4" if args.scale == 1.0 else f"C4 x {args.scale} (debug, not a bench value)",
                       "nq": nq, "ns": ns, "nf": nf, "nt": nt, "y_density": d["y_density"], "alpha": d["alpha"],
                       "weighted": d["weighted"], "clean_fused": True,
                       "sharding": "single GPU" if world == 1 else
                       (f"query rows x{world}; T by target-column block, " +
                        ("blocks stored to all peers from the T-GEMM epilogue over NVLink P2P (fused all-gather)"
                         if sharded.fused else "NCCL all-gather") +
                        f"; collectives behind the C ABI (ss_comm_*, NCCL {comm.nccl_version()} dlopen()ed by the library)"),
                       "l2": "inputs (27 GB) and output (40 GB) exceed the 126 MB L2; no flush needed"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "check": checkres, "opt_in_f64_int8": None,
        }

    # ---- opt-in FP64-grade mode on the INT8 tensor pipe, measured beside the headline (untimed for `value`).
    # A watchdog prints the line without it if this side measurement does not finish: an opt-in mode must never
    # cost the headline (the watchdog thread runs while the main thread sits in the library call).
    if world == 1 and not pflag and not args.no_check:
        import threading

        def _bail():
            line["opt_in_f64_int8"] = {"error": "the opt-in INT8 measurement did not finish within 240 s; not reported"}
            print(json.dumps(line), flush=True)
            os._exit(0)

        dog = threading.Timer(240.0, _bail)
        dog.daemon = True
        dog.start()
        try:
            flag8 = SS_PREDICT_CLEAN | (3 << 4)
            ref_sample = bR[tt, tq].clone()  # FP64 DMMA result at the sampled entries
            check(L.ss_predict_query(ctx.h, mXq.h, mXs.h, mY.h, mR.h, flag8, None))  # warm-up (allocates slices)
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(ext)
            check(L.ss_predict_query(ctx.h, mXq.h, mXs.h, mY.h, mR.h, flag8, None))
            f1.record(ext)
            barrier()
            ms8 = f0.elapsed_time(f1)
            st8 = ctx.int8_stats()
            got8 = bR[tt, tq]
            rel8 = ((got8 - want).abs() / want.abs().clamp_min(1e-300)).max().item()
            line["opt_in_f64_int8"] = {
                "mode": "precision=f64_int8: 6 unsigned 8-bit slices per operand, exact INT32 slice products on tcgen05 "
                        "kind::i8 (planes of zeros skipped: 6 pairs for the 0/1 label matrix, 21 for R = Xq*T), FP64 "
                        "recombination; every entry certified a posteriori (error bound <= 4e-13 of the entry per "
                        "product), a product that fails is re-run on the FP64 DMMA path (opt-in)",
                "products_on_int8_pipe": st8[0], "products_rerun_in_fp64": st8[1], "uncertified_entries_last_product": st8[2],
                "value": nq * nt / (ms8 * 1e-3), "unit": UNIT, "ms_per_step": ms8,
                "speedup_vs_fp64_dmma": ms_step / ms8, "max_rel_err_sampled": rel8,
                "max_rel_diff_vs_dmma_sampled": ((got8 - ref_sample).abs() / ref_sample.abs().clamp_min(1e-300)).max().item()}
            check(L.ss_predict_query(ctx.h, mXq.h, mXs.h, mY.h, mR.h, SS_PREDICT_CLEAN, None))  # restore the FP64 result
        except Exception as exc:  # noqa: BLE001 -- the side measurement may fail, the headline may not
            line["opt_in_f64_int8"] = {"error": f"{type(exc).__name__}: {exc}"}
        dog.cancel()

    # ---- the call a user makes (NamedArrays in pageable memory) and the other BASELINE configs, beside the headline
    if world == 1 and not pflag and not args.no_side:
        import threading

        def _bail2():
            line.setdefault("side_configs", {"error": "the side measurements did not finish within 420 s; not reported"})
            print(json.dumps(line), flush=True)
            os._exit(0)

        dog = threading.Timer(420.0, _bail2)
        dog.daemon = True
        dog.start()
        del bXq, bR, mXq, mR
        torch.cuda.empty_cache()
        try:
            line["e2e_api"] = run_e2e_api(ss, np, d)
        except Exception as exc:  # noqa: BLE001
            line["e2e_api"] = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            del bXs, bY, mXs, mY
            torch.cuda.empty_cache()
            line["side_configs"] = run_side_configs(ss, np, torch, ctx, dev)
        except Exception as exc:  # noqa: BLE001
            line["side_configs"] = {"error": f"{type(exc).__name__}: {exc}"}
        dog.cancel()

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        sharded.close()
        comm.close()
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e_api(ss, np, d, nq_api=10_000):
    """The call a user makes: construct((ytrain, ytest), (Xtrain, Xtest)) + predict((A, B), ytest, clean=True) on
    NamedArrays that live in ordinary (pageable) NumPy memory, result back as a NamedArray; C4's inner dimensions with
    `nq_api` queries.  Beside it the C entry point (`ss_predict_query_host`, pinned buffers) on the same shape."""
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    ns, nf, nt = d["ns"], d["nf"], d["nt"]
    rng = np.random.default_rng(SEED + 7)
    Xs = np.asfortranarray(np.round(rng.random((ns, nf)), 6))
    Xq = np.asfortranarray(np.round(rng.random((nq_api, nf)), 6))
    Y = np.asfortranarray((rng.random((ns, nt)) < d["y_density"]).astype(np.float64))
    Yq = np.zeros((nq_api, nt), order="F")
    sn, qn = [f"s{i}" for i in range(ns)], [f"q{i}" for i in range(nq_api)]
    fn, tn = [f"f{i}" for i in range(nf)], [f"t{i}" for i in range(nt)]
    ytrain, ytest = ss.NamedArray(Y, (sn, tn)), ss.NamedArray(Yq, (qn, tn))
    Xtrain, Xtest = ss.NamedArray(Xs, (sn, fn)), ss.NamedArray(Xq, (qn, fn))

    def api_step():
        A, B = ss.construct((ytrain, ytest), (Xtrain, Xtest))
        return ss.predict((A, B), ytest, clean=True, layout="dense")

    got = api_step()  # warm-up
    t0 = time.perf_counter()
    got = api_step()
    t_api = time.perf_counter() - t0
    # the C entry point on the same shape, pinned buffers
    ctx, L = ss.Context.default(), ss.lib()
    R = np.empty((nq_api, nt), order="F")
    ptrs = []
    for a in (Xq, Xs, Y, R):
        p = C.c_void_p()
        check(L.ss_host_alloc(a.size * 8, C.byref(p)))
        ptrs.append(p)
        if a is not R:
            C.memmove(p, a.ctypes.data, a.size * 8)
    pXq, pXs, pY, pR = ptrs
    call = lambda: check(L.ss_predict_query_host(ctx.h, pXq, nq_api, pXs, ns, pY, ns, nq_api, ns, nf, nt, SS_PREDICT_CLEAN, pR, nq_api))
    call()
    t0 = time.perf_counter()
    call()
    t_entry = time.perf_counter() - t0
    Rn = np.ctypeslib.as_array(C.cast(pR, C.POINTER(C.c_double)), shape=(nt, nq_api)).T
    same = bool(np.array_equal(got.array, Rn))
    for p in ptrs:
        L.ss_host_free(p)
    return {"value": nq_api * nt / t_api, "unit": UNIT, "queries": nq_api, "ms": t_api * 1e3,
            "h2d_bytes": int((Xq.size + Xs.size + Y.size) * 8), "d2h_bytes": int(nq_api * nt * 8),
            "api": "construct((ytrain, ytest), (Xtrain, Xtest)) + predict((A, B), ytest, clean=True) on pageable NumPy NamedArrays",
            "entry_point_same_shape": {"value": nq_api * nt / t_entry, "ms": t_entry * 1e3, "api": "ss_predict_query_host, pinned buffers"},
            "ratio_api_over_entry_point": t_entry / t_api, "bit_identical_to_entry_point": same}


def run_side_configs(ss, np, torch, ctx, dev):
    """The other BASELINE configs on this GPU, each with its own check against the oracle (untimed): C2 = Enzyme-shaped
    10-fold CV, C3 = 21-point weighted alpha sweep, C5 = sparse recommender (5 % of the users of the 2M x 500k graph).
    Reported beside the headline; not part of `value`."""
    import ctypes
    from oracle import simspread_oracle as o
    from simspread_b200._lib import check
    out = {}
    L = ss.lib()
    # ---- C2 -----------------------------------------------------------------------------------------------------
    S, Y, names, tn, folds, alpha, weighted = synth_small_config(np, "C2")
    DD, DT = ss.NamedArray(S, (names, names)), ss.NamedArray(Y, (names, tn))
    ss.cross_validate(DT, DD, alpha, weighted=weighted, folds=folds)
    ts, launches, tf = [], 0, []
    for _ in range(5):
        l0 = ctx.launch_count()
        tm = {}
        t0 = time.perf_counter()
        res = ss.cross_validate(DT, DD, alpha, weighted=weighted, folds=folds, timing=tm)
        ts.append(time.perf_counter() - t0)
        launches = ctx.launch_count() - l0
        tf.append(tm)
    Xo, xr, xc = o.featurize(S, names, names, alpha, weighted)
    worst, row = 0.0, 0
    for q in folds:
        qi = [names.index(x) for x in q]
        si = [i for i in range(len(names)) if names[i] not in set(q)]
        Xq_, Xs_, Y_ = Xo[np.ix_(qi, si)], Xo[np.ix_(si, si)], Y[si]
        want = o.predict_blocks_query(Xq_, Xs_, Y_)
        o.clean_blocks(want, o.degrees_blocks(Xs_, Y_)[2])
        got = res["yhat"].array[row:row + len(q)]
        row += len(q)
        nz = want != 0
        if nz.any():
            worst = max(worst, float(np.max(np.abs(got[nz] - want[nz]) / np.abs(want[nz]))))
        if not np.array_equal(got[~nz], want[~nz]):
            worst = float("inf")
    yb = np.concatenate([Y[[names.index(x) for x in q]] for q in folds]).ravel() > 0
    out["C2_enzyme_10fold_cv"] = {
        "shape": [len(names), len(tn)], "alpha": alpha, "weighted": weighted, "ms_per_cv_wall_median_of_5": float(np.median(ts)) * 1e3,
        "scores_per_s": len(names) * len(tn) / float(np.median(ts)), "kernel_launches_per_cv": int(launches),
        "folds_call_ms_median": float(np.median([t_["folds_call_ms"] for t_ in tf])),
        "folds_call_kernel_launches": int(tf[-1]["folds_call_launches"]),
        "includes": "upload, featurize, 10 folds (gather, degrees, spread, T, R + clean!), AuROC/AuPRC, @20 metrics, download",
        "check": {"max_rel_err_vs_oracle_block_form": worst, "tolerance": 1e-12,
                  "AuROC_rel_diff_vs_oracle": abs(res["AuROC"] - o.AuROC(yb, res["yhat"].array.ravel())) / max(res["AuROC"], 1e-300),
                  "ok": bool(worst < 1e-12)}}
    # ---- C3 -----------------------------------------------------------------------------------------------------
    S, Y, names, tn, folds, _, weighted = synth_small_config(np, "C3")
    DD, DT = ss.NamedArray(S, (names, names)), ss.NamedArray(Y, (names, tn))
    q = folds[0]
    alphas = [round(0.05 * i, 2) for i in range(21)]
    ss.alpha_sweep(DT, DD, q, alphas[:2] + alphas[-1:])  # warm-up: the dense and the sparse chain allocate their workspaces
    runs = []
    for _ in range(3):
        tm_ = {}
        t0 = time.perf_counter()
        sw = ss.alpha_sweep(DT, DD, q, alphas, timing=tm_)
        runs.append((time.perf_counter() - t0, tm_))
    t_sw, tm = sorted(runs, key=lambda r: r[0])[1]  # median of 3 complete sweeps
    a_chk = 0.5
    nq = len(q)
    Xo = o.cutoff(S[:, nq:], a_chk, True)
    want = o.predict_blocks_query(Xo[:nq], Xo[nq:], Y[nq:])
    o.clean_blocks(want, o.degrees_blocks(Xo[nq:], Y[nq:])[2])
    au_want = o.AuROC(Y[:nq].ravel() > 0, want.ravel())
    pt = [p_ for p_ in sw if abs(p_["alpha"] - a_chk) < 1e-9][0]
    out["C3_alpha_sweep_21_points"] = {
        "shape": {"nq": nq, "ns": len(names) - nq, "nt": len(tn)}, "wall_s": t_sw, "wall_s_all_runs": [r[0] for r in runs],
        "timing": "median of 3 complete 21-point sweeps after one warm-up sweep of 3 points", "setup_s": tm.get("setup_s"),
        "sweep_s": tm.get("sweep_s"), "ms_per_alpha": (tm.get("sweep_s") or t_sw) / 21 * 1e3, "scores_per_s": 21 * nq * len(tn) / t_sw,
        "layouts": [p_.get("layout") for p_ in sw],
        "check": {"alpha": a_chk, "AuROC": pt["AuROC"], "AuROC_oracle_block_form": au_want,
                  "rel_diff": abs(pt["AuROC"] - au_want) / au_want, "tolerance": 1e-12, "ok": bool(abs(pt["AuROC"] - au_want) <= 1e-12 * au_want)}}
    del DD, DT, S, Y
    # ---- C5 -----------------------------------------------------------------------------------------------------
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from c5_graph import make_graph, partial_products, wrap
    import scipy.sparse as sp
    ns, nt, topl, frac = 2_000_000, 500_000, 20, 0.05
    G = make_graph(ns, nt, 1e-4, dev, "poisson", False)
    hY, hYT = wrap(L, ctx, check, G)
    idx = torch.full((ns, topl), -2, dtype=torch.int32, device=dev)
    val = torch.zeros((ns, topl), dtype=torch.float64, device=dev)
    vi, vm = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.ss_ivec_wrap(ctx.h, ctypes.c_void_p(idx.data_ptr()), ns * topl, ctypes.byref(vi)))
    check(L.ss_mat_wrap(ctx.h, ctypes.c_void_p(val.data_ptr()), topl, ns, topl, ctypes.byref(vm)))
    ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    hU = ctypes.c_void_p()
    e0.record(ext)
    check(L.ss_transfer_build(ctx.h, hY, hYT, ctypes.byref(hU)))
    e1.record(ext)
    users = int(ns * frac)
    check(L.ss_recommend_topl_transfer(ctx.h, hY, hU, topl, 0, 4096, vi, vm))  # warm-up
    ctx.sync()
    e1.record(ext)
    check(L.ss_recommend_topl_transfer(ctx.h, hY, hU, topl, 0, users, vi, vm))
    e2.record(ext)
    ctx.sync()
    ms_stream = e1.elapsed_time(e2)
    info = (ctypes.c_int64 * 4)()
    check(L.ss_transfer_info(hU, info))
    first = (idx[:users].clone(), val[:users].clone())
    check(L.ss_recommend_topl_transfer(ctx.h, hY, hU, topl, 0, users, vi, vm))
    ctx.sync()
    identical = bool(torch.equal(first[0], idx[:users]) and torch.equal(first[1].view(torch.int64), val[:users].view(torch.int64)))
    check(L.ss_transfer_destroy(hU))
    pp = partial_products(G, users)
    Ysp = sp.csr_matrix((np.ones(G["nnz"]), G["y_idx"].cpu().numpy(), G["y_ptr"].cpu().numpy()), shape=(ns, nt))
    rows = np.sort(np.random.default_rng(5).choice(users, size=6, replace=False))
    F, _ = o.two_layer_scores_sparse(Ysp, rows)
    gi, gv = first[0].cpu().numpy(), first[1].cpu().numpy()
    exact = True
    for j, r in enumerate(rows):
        dense = np.zeros(nt)
        sl = slice(F.indptr[j], F.indptr[j + 1])
        dense[F.indices[sl]] = F.data[sl]
        order = o.sortperm_rev(dense)[:topl]
        exact = exact and np.array_equal(gi[r], order) and np.array_equal(gv[r].view(np.uint64), dense[order].view(np.uint64))
    out["C5_sparse_recommender_5pct_users"] = {
        "users": ns, "items": nt, "edges": int(G["nnz"]), "L": topl, "users_ranked": users, "ms": ms_stream,
        "partial_products": pp, "partial_products_per_s": pp / (ms_stream * 1e-3), "scores_per_s": users * nt / (ms_stream * 1e-3),
        "algorithmic_gbs_10B_per_partial_product": pp * 10 / (ms_stream * 1e-3) / 1e9,
        "transfer_matrix": {"entries": int(info[0]), "bytes": int(info[1]), "tile_width": int(info[2])},
        "kernel": "tr_stream_kernel (shared-memory accumulators, no atomics) over the materialised transfer matrix",
        "check": {"sampled_users": int(len(rows)), "top20_order_and_scores_bit_identical_to_scipy_sparse": bool(exact),
                  "second_run_bit_identical": identical, "ok": bool(exact and identical)}}
    return out


def host_mem_available():
    try:
        for l in open("/proc/meminfo"):
            if l.startswith("MemAvailable:"):
                return int(l.split()[1]) * 1024
    except OSError:
        pass
    return 0


def run_e2e(args, ss, ctx, torch, comm, np, dev, world, rank, d, nq_l, nt_l, bXq, bXs, bY, ldq, lds, ldy, sharded,
            mats):
    """Same metric through the host-buffer entry points: pinned host inputs are copied to the device and R is copied
    back inside the timed region, every step.  N = 1: ss_predict_query_host.  N > 1, per rank and per step: upload of
    this rank's column shard of the replicated Xs (1/N of it; one NCCL all-gather over NVLink assembles Xs on every
    GPU instead of N copies of the whole matrix over PCIe), upload of its Y block, the sharded front
    (ss_sharded_front), and its query slab streamed through ss_stream_product_host (H2D / GEMM / D2H pipelined)."""
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    L = ss.lib()
    nq, ns, nf, nt = d["nq"], d["ns"], d["nf"], d["nt"]
    mXq, mXs, mY, mR, bR, ldr = mats
    xs_cols = nf // world if nf % world == 0 else nf  # Xs shard of this rank (whole matrix if it does not divide)
    xs_c0 = rank * xs_cols if xs_cols != nf else 0
    # host memory budget: shrink the query slab if the box cannot pin everything
    fixed = (ns * xs_cols + ns * nt_l) * 8
    per_q = (nf + nt) * 8
    avail = host_mem_available() // max(1, world)
    nq_e = nq_l
    if avail and fixed + nq_e * per_q > 0.6 * avail:
        nq_e = max(128, int((0.6 * avail - fixed) // per_q) // 128 * 128)
    if world > 1 and nq_e != nq_l:
        return {"value": None, "unit": UNIT, "note": "not enough host memory to pin the inputs"}

    def pinned(rows, cols):
        p = C.c_void_p()
        check(L.ss_host_alloc(rows * cols * 8, C.byref(p)))
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(cols, rows))  # column-major
        return p, arr

    pXq, hXq = pinned(nq_e, nf)
    pXs, hXs = pinned(ns, xs_cols)
    pY, hY = pinned(ns, nt_l)
    pR, hR = pinned(nq_e, nt)
    # fill the host buffers with the device-resident operands (untimed)
    torch.cuda.synchronize()
    torch.from_numpy(hXq).copy_(bXq[:, :nq_e])
    torch.from_numpy(hXs).copy_(bXs[xs_c0:xs_c0 + xs_cols, :ns])
    torch.from_numpy(hY).copy_(bY[:, :ns])
    torch.cuda.synchronize()

    if world == 1:
        def e2e_step():
            check(L.ss_predict_query_host(ctx.h, pXq, nq_e, pXs, ns, pY, ns, nq_e, ns, nf, nt, SS_PREDICT_CLEAN,
                                          pR, nq_e))
    else:
        hT, hkt = sharded.views()

        def e2e_step():
            check(L.ss_mat_upload_cols_async(ctx.h, mXs.h, xs_c0, xs_cols, pXs, ns))
            if xs_cols != nf:
                check(L.ss_comm_allgather_cols(comm.h, mXs.h))
            check(L.ss_mat_upload(ctx.h, mY.h, pY, ns))
            sharded.front(mXs, mY)  # degrees, two NCCL collectives, T tiles stored into every rank's T
            check(L.ss_stream_product_host(ctx.h, pXq, nq_e, nq_e, hT, hkt, pR, nq_e))

    e2e_step()  # warm-up (allocates the streaming workspaces)
    ctx.sync()
    if world > 1:
        comm.barrier()
    nsteps = max(1, min(args.steps, 2))
    # the call is synchronous and spans three streams (H2D / compute / D2H): the clock around the
    # synchronous call is the honest end-to-end time
    t0 = time.perf_counter()
    for _ in range(nsteps):
        e2e_step()
    ctx.sync()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / nsteps
    if world > 1:
        wall = comm.allreduce_host([wall], "max")[0]
    got = torch.from_numpy(hR)[:64, :64]
    same = bool(torch.equal(got.to(dev), bR[:64, :64]))
    if world > 1:
        same = bool(comm.allreduce_host([0.0 if same else 1.0], "max")[0] == 0.0)
    res = {
        "value": (nq_e * world) * nt / wall, "unit": UNIT,
        "h2d_bytes_per_step": int((nq_e * nf + ns * xs_cols + ns * nt_l) * 8 * world),
        "d2h_bytes_per_step": int(nq_e * nt * 8 * world),
        "ms_per_step": wall * 1e3, "steps": nsteps, "queries": nq_e * world,
        "api": "ss_predict_query_host (pinned host buffers, slab-pipelined H2D / GEMM / D2H)" if world == 1 else
               "per rank: ss_mat_upload_cols_async(Xs shard) + ss_comm_allgather_cols (NVLink) + ss_mat_upload(Y block) + "
               "ss_sharded_front (T over NVLink) + ss_stream_product_host",
        "matches_resident_run": same,
    }
    for p in (pXq, pXs, pY, pR):
        L.ss_host_free(p)
    return res


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
