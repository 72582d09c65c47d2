#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

Workload (N=1): config 3 of BASELINE.json -- one open TensorTrain, L=256, q=2, d=256, Float64,
`rand_tt` fill (U[0,1), NumPy Philox seed 3+rank), full `compress!(TruncThresh(1e-10))` =
eps=0 right sweep + truncating left sweep = 2(L-1) = 510 site-steps.  A bench "step" is one whole
compress! on a fresh copy of the train.  metric = compress! site-steps/s.
N>1: a single-train sweep is sequential in L and does not shard (DESIGN.md "replicas only"), so every
rank compresses its own train (seed 3+rank): weak scaling, no data-path collective.

JSON line keys beyond the base contract:
  roofline     dominant kernel of the step, algorithmic FLOPs per launch / CUDA-event time per launch
               against an FP64 peak measured in this run (cuBLAS DGEMM through torch.matmul); traffic from
               profiles/ncu_traffic.json (the committed `ncu --set full` summaries)
  cpu_baseline the CPU oracle (NumPy/SciPy-OpenBLAS gesdd restatement of the reference; Julia is not
               installed) on the same workload, host cores of this box
  parity_in_run  retained bond dimensions of the timed compress! against the CPU arm's on the same input
  extra        every other BASELINE.json configuration, each with its own CPU value and in-run parity check:
               sample (config 4 at 10^8 draws, strong-scaled over the ranks), batch (config 5: 4096 periodic
               trains sharded over the ranks, pair marginals all-gathered between device buffers), config2,
               randn_fill (config 3 with the full-rank fill, L = 256), pairs_of_one_train (the third sharding
               axis of SURVEY 8e).  At N > 1 rank 0 repeats every sharded leg alone in the same job and the line
               carries speedup_vs_1gpu next to an identity check of the results.
A failed parity / identity check is recorded in the line and fails the run (exit code) AFTER the line is printed.
`--impl reference` times the CPU oracle (the reference's algorithm; kind "port") on a bounded sample
(`--fill randn` = the full-rank fill of config 3, reported as extra.randn_fill of the main arm); when the sample
is the full chain its line carries the retained bond dimensions, and the main arm asserts the CUDA path's are
identical (outside the timed region).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tensortrains.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

L3, D3, Q3 = 256, 256, 2
L4, D4, Q4 = 256, 128, 4
EPS3 = 1e-10


def make_open(seed, L, d, q, fill="rand"):
    rng = np.random.Generator(np.random.Philox(seed))
    bonds = [1] + [d] * (L - 1) + [1]
    if fill == "rand":
        return bonds, [rng.random((bonds[t], bonds[t + 1], q)) for t in range(L)]
    s = 1.0 / math.sqrt(d)
    return bonds, [s * rng.standard_normal((bonds[t], bonds[t + 1], q)) for t in range(L)]


def pack(tensors):
    return np.concatenate([np.reshape(a, -1, order="F") for a in tensors])


# ---- algorithmic work ----------------------------------------------------
def qr_sweep_flops(bonds, Q, bw=32):
    """eps=0 right sweep over an open train: per-kernel-class algorithmic FLOPs and launch counts.
    Factor 2 m^2 (n - m/3) split into panel + trailing update, thin-Q formation, absorb 2 a m m' Q."""
    L = len(bonds) - 1
    b = list(bonds)
    out = {"k_qr_panel": [0, 0.0], "k_qr_update": [0, 0.0], "k_apply_q": [0, 0.0], "k_absorb": [0, 0.0],
           "k_qr_stream": [0, 0.0]}
    for t in range(L - 1, 0, -1):
        p, n, a = b[t + 1] * Q, b[t], b[t - 1]
        k = min(p, n)
        nblk = (n + 31) // 32
        if p >= n and 2 <= nblk <= 8 and p <= 512:
            # the whole site QR is ONE cluster kernel (qr_stream.cuh): panels + trailing updates
            out["k_qr_stream"][0] += 1
            out["k_qr_stream"][1] += 2.0 * k * k * (p - k / 3.0)
            out["k_apply_q"][0] += 1
            out["k_apply_q"][1] += 2.0 * k * k * (p - k / 3.0)
            out["k_absorb"][0] += 1
            out["k_absorb"][1] += 2.0 * a * n * k * Q
            b[t] = k
            continue
        for j0 in range(0, k, bw):
            bwc, rows = min(bw, k - j0), p - j0
            out["k_qr_panel"][0] += 1
            out["k_qr_panel"][1] += 2.0 * bwc * bwc * (rows - bwc / 3.0)
            trail = n - j0 - bwc
            if n - j0 - 1 > 0:
                out["k_qr_update"][0] += 1
                out["k_qr_update"][1] += 4.0 * rows * bwc * max(trail, 0)
        out["k_apply_q"][0] += 1
        out["k_apply_q"][1] += 2.0 * k * k * (p - k / 3.0)
        out["k_absorb"][0] += 1
        out["k_absorb"][1] += 2.0 * a * n * k * Q
        b[t] = k
    return out, b


def clocks_sampler(stop, out, dev):
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active,"
                              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                              "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
    except Exception:
        return
    def rd():
        for line in p.stdout:
            out.append(line.strip())
    th = threading.Thread(target=rd, daemon=True)
    th.start()
    stop.wait()
    p.terminate()


def summarize_clocks(lines):
    sm, mx, reasons = [], 0.0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        try:
            sm.append(float(f[0])); mx = max(mx, float(f[1]))
        except Exception:
            continue
        for nm, v in zip(names, f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(nm)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(torch, dev):
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


LAST_CPU_BONDS = None


def cpu_compress_steps_per_s(L, d, q, seed, reps=1, fill="rand"):
    global LAST_CPU_BONDS
    from oracle import tt_oracle as O
    _, ts = make_open(seed, L, d, q, fill)
    best = 1e99
    for _ in range(reps):
        A = O.TensorTrain([a.copy() for a in ts])
        t0 = time.perf_counter()
        O.compress(A, O.TruncThresh(EPS3))
        best = min(best, time.perf_counter() - t0)
        LAST_CPU_BONDS = [int(x) for x in O.bond_dims(A)]
    return 2 * (L - 1) / best, best


def cpu_sample_draws_per_s(L, d, q, seed, n):
    from oracle import tt_oracle as O
    _, ts = make_open(seed, L, d, q)
    A = O.TensorTrain(ts)
    O.normalize(A)
    rz = O.accumulate_R(A)
    rng = np.random.default_rng(0)
    t0 = time.perf_counter()
    for _ in range(n):
        us = rng.random(L)
        O.sample(lambda t: us[t], A, rz=rz)
    return n / (time.perf_counter() - t0)


def reference_arm(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; Julia+MKL cannot run here).
    A step is the config-3 compress! itself (510 site-steps) when the whole --steps/--warmup run fits ~2.5
    minutes on this box, otherwise an L=64 or L=32 slice of the same chain (same 256 x 512 bulk matrices; the
    ends of a chain are cheap, so a slice OVERSTATES the CPU's site-steps/s -- said in `sample`).  The choice is
    made from one probe run of the L=32 slice."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fill = args.fill
    # BLAS thread count: all host threads is NOT the fastest setting for 256 x 512 gesdd calls (measured: 8 of 8
    # threads 4.5x slower than 4 in this image); probe the L=32 slice at 1, 2, 4, ... threads and keep the best
    ncpu = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    cands = sorted({n for n in (1, 2, 4, 8, 16, 32, ncpu) if n <= ncpu}) if threadpool_limits else [ncpu]
    probes = {}
    for n in cands:
        ctx = threadpool_limits(limits=n) if threadpool_limits else None
        t0 = time.perf_counter()
        if ctx:
            with ctx:
                cpu_compress_steps_per_s(32, D3, Q3, 3, fill=fill)
        else:
            cpu_compress_steps_per_s(32, D3, Q3, 3, fill=fill)
        probes[n] = time.perf_counter() - t0
    nthr = min(probes, key=probes.get)
    probe = probes[nthr]
    if threadpool_limits:
        _keep = threadpool_limits(limits=nthr)   # stays in force for the rest of the process
    nrun = max(args.steps + args.warmup, 1)
    # bulk right-sweep factorisations dominate: 23 of them in the L=32 slice, 55 at L=64, 247 in the full chain
    if probe * (247.0 / 23.0) * nrun <= 150.0:
        Ls = L3
    elif probe * (55.0 / 23.0) * nrun <= 150.0:
        Ls = 64
    else:
        Ls = 32
    for _ in range(args.warmup):
        cpu_compress_steps_per_s(Ls, D3, Q3, 3, fill=fill)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_compress_steps_per_s(Ls, D3, Q3, 3, fill=fill)
    dt = time.perf_counter() - t0
    # the timed region includes input generation (~5%); subtract nothing, say so
    val = args.steps * 2 * (Ls - 1) / dt
    what = ("the full C3 chain (L=256, 510 site-steps) per step" if Ls == L3 else
            f"L={Ls} slice of the L=256 chain per step ({2 * (Ls - 1)} of 510 site-steps, same 256x512 bulk matrices; "
            "the cheap chain ends weigh more in a slice, which overstates the CPU rate)")
    line = {"impl": "reference", "metric": "compress_site_steps_per_s", "value": val, "unit": "site-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3 single TT q=2 d=256 {fill}_tt compress!(TruncThresh(1e-10))", "sample": what, "fill": fill,
                       "bonds_after": LAST_CPU_BONDS if Ls == L3 else None, "seed": 3,
                       "probe_L32_s": probe, "blas_threads": nthr, "host_cores": ncpu,
                       "probe_L32_s_by_threads": {str(k): round(v, 3) for k, v in probes.items()}},
            "cpu_baseline": {"value": val, "unit": "site-steps/s", "cores": nthr, "kind": "port",
                             "sample": what + f"; NumPy/SciPy-OpenBLAS gesdd restatement, not Julia/MKL; BLAS threads = {nthr} of {ncpu} host "
                                                f"cores (the fastest of {cands} on a probe); includes input generation"},
            "e2e": {"value": val, "unit": "site-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-extra", action="store_true", help="skip the sampling / randn / cpu legs")
    ap.add_argument("--fill", default="rand", choices=["rand", "randn"], help="--impl reference only: fill of the C3 train")
    ap.add_argument("--draws", type=int, default=int(os.environ.get("TTB_BENCH_DRAWS", "100000000")),
                    help="total draws of the config-4 leg (BASELINE.json: 10^8), sharded over the ranks")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import ttb200 as T

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    T.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its version banner to stdout when the communicator is created; rank 0 must print ONE
        # JSON line, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K, W = args.steps, max(args.warmup, 0)
    bonds, ts = make_open(3 + rank, L3, D3, Q3)
    nbytes = sum(a.size for a in ts) * 8
    pin = T.PinnedBuffer(nbytes // 8)
    pin.array[:] = pack(ts)
    del ts
    tr = T.TruncThresh(EPS3)
    nsteps_unit = 2 * (L3 - 1)

    # ---- device-resident arm: fresh copies made outside the timed region ----
    base = T.TensorTrain.empty(bonds, (Q3,))
    base.upload_all(pin.array)
    copies = [base.copy() for _ in range(W + K)]
    stop, clk = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, clk, local), daemon=True)
    th.start()
    for i in range(W):
        T.compress_(copies[i], tr)
    barrier()
    dev_ms, launches, ms_each = 0.0, 0, []
    w0 = time.perf_counter()
    for i in range(W, W + K):
        T.compress_(copies[i], tr)
        ms, nl = copies[i].last_timing()
        dev_ms += ms; launches += nl; ms_each.append(round(ms, 3))
    barrier()
    wall = time.perf_counter() - w0
    dev_ms = max_over_ranks(dev_ms)
    bonds_after = T.bond_dims(copies[-1])
    out_params = T.nparams(copies[-1])

    # ---- end-to-end arm: pinned host -> device -> compress! -> host, every step ----
    e2e_h = T.TensorTrain.empty(bonds, (Q3,))
    outbuf = T.PinnedBuffer(nbytes // 8)
    def e2e_step():
        # one long-lived handle per worker, like a caller compressing a stream of trains: reset the bond
        # dimensions, upload asynchronously (sites last-to-first: the right sweep starts on the first
        # chunk), compress, read the compressed train back
        e2e_h.reset()
        e2e_h.upload_all(pin.array, async_=True)
        T.compress_(e2e_h, tr)
        n = T.nparams(e2e_h)
        e2e_h.download_all(outbuf.array[:n])
        return n
    e2e_step()
    barrier()
    e0 = time.perf_counter()
    for _ in range(K):
        nout = e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
    stop.set()
    del e2e_h

    value = world * K * nsteps_unit / (dev_ms * 1e-3)
    e2e_val = world * K * nsteps_unit / e2e_s

    line = {"metric": "compress_site_steps_per_s", "value": value, "unit": "site-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3: single TensorTrain L=256 q=2 d=256 Float64, rand_tt fill, compress!(TruncThresh(1e-10)) "
                                   "= 510 site-steps per step; one train per GPU (replicas, no collective)",
                       "l2": "input 254 MiB per train > 126 MB L2; every step runs on a fresh copy",
                       "timing": "CUDA events on the library stream around each compress! call, summed over steps, max over ranks",
                       "wall_s": wall, "ms_each": ms_each, "bonds_after_max": int(max(bonds_after))},
            "e2e": {"value": e2e_val, "unit": "site-steps/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(nout) * 8,
                    "note": "pinned host tensors -> ttb_reset + ttb_upload_all_async -> ttb_compress -> ttb_download_all, wall clock"},
            "gpu_launches": int(launches),
            "clocks": summarize_clocks(clk)}

    pre_fail = []      # identity checks of sharded legs against rank 0's own one-GPU run (messages of the ones that failed)

    def alone_on_rank0(fn):
        """At N > 1: rank 0 repeats a leg ALONE on the whole workload (the other ranks wait at the barrier), so the
        line carries the 1-GPU figure measured in the same job and speedup_vs_1gpu is arithmetic on two numbers of
        this run.  Returns fn()'s result on rank 0, None elsewhere."""
        res = None
        barrier()
        if rank == 0:
            res = fn()
        barrier()
        return res

    # ---- sampling half of the metric (config 4): the train is replicated, the --draws (10^8) draws are sharded by
    #      global draw index (Philox counter) over the ranks = STRONG scaling; chunked launches, histograms stay on
    #      the device per chunk, the only collective is the all-reduce of the L x Q histogram ----
    sample_info = None
    if not args.no_extra:
        from ttb200 import parallel as P
        _, t4 = make_open(4, L4, D4, Q4)
        A4 = T.TensorTrain(t4)
        del t4
        T.normalize_(A4)
        rz4 = T.accumulate_R(A4, keep=True)[0]            # rz= keyword: the environment pass is paid once, not per chunk
        ND = int(args.draws)
        CH = 10_000_000

        def draw_range(lo, hi):
            hist = np.zeros((L4, Q4), dtype=np.int64)
            ms, slp = 0.0, 0.0
            t0 = time.perf_counter()
            for a in range(lo, hi, CH):
                h, s_ = T.sample_stats(A4, min(CH, hi - a), seed=1, first_index=a, rz=rz4)
                hist += h; slp += s_
                ms += A4.last_timing()[0]
            return hist, slp, ms, time.perf_counter() - t0

        T.sample_stats(A4, 20000, seed=1, first_index=0, rz=rz4)   # warm-up (pack buffers, attributes)
        lo4, hi4 = P.shard_range(ND, rank, world)
        barrier()
        hist, slp, ms4, wall4 = draw_range(lo4, hi4)
        ms4 = max_over_ranks(ms4)
        wall4 = max_over_ranks(wall4)
        hist_tot = P.allreduce_sum(hist)
        assert int(hist_tot.sum()) == ND * L4
        sample_info = {"workload": f"C4: TensorTrain L=256 q=4 d=128 rand_tt, normalize! then {ND:.0e} draws in all (BASELINE.json "
                                   "configs[3]), sharded over the ranks by global draw index (strong scaling), chunks of 1e7",
                       "draws": ND, "draws_per_rank": hi4 - lo4, "draws_per_s": ND / (ms4 * 1e-3), "device_ms": ms4,
                       "e2e_draws_per_s": ND / wall4, "alg_mflop_per_draw": 8.65,
                       "hist_checksum": int((hist_tot * np.arange(1, Q4 + 1)).sum()),
                       "note": "device time of ttb_sample_stats (pack + tiled DMMA sampler; accumulate_R passed as rz=); the "
                               "histogram (L x Q int64) is the only result copied back; hist_checksum is independent of N "
                               "(Philox counter = global draw index)"}
        if world > 1:
            one = alone_on_rank0(lambda: draw_range(0, ND))
            if rank == 0:
                h1, _, ms1, wall1 = one
                same4 = bool(np.array_equal(h1, hist_tot))       # a mismatch fails the run AFTER the line is printed
                pre_fail.append(None if same4 else "sharded C4 histogram differs from the one-GPU run")
                sample_info["one_gpu"] = {"draws_per_s": ND / (ms1 * 1e-3), "e2e_draws_per_s": ND / wall1, "device_ms": ms1}
                sample_info["speedup_vs_1gpu"] = {"device": ms1 / ms4, "e2e": wall1 / wall4,
                                                  "histogram_identical_to_one_gpu_run": same4}
        del A4, rz4

    # ---- config 5: 4096 periodic trains (64 sites, bond 32, q=2) sharded over the ranks by contiguous
    #      train ranges (strong scaling of this leg); normalize! + twovar_marginals + compress!(1e-6);
    #      collectives (NCCL): all-gather of per-train log Z, bond dimensions and of the pair marginals, the
    #      latter between DEVICE buffers (258 MiB in all), then one copy into a pinned host buffer ----
    batch_info = None
    if not args.no_extra:
        import ctypes as C
        from ttb200 import parallel as P
        B5, L5, D5, Q5 = 4096, 64, 32, 2
        site5 = D5 * D5 * Q5
        npair = C.c_int64()

        def gen5(lo, hi):
            packed = np.empty((hi - lo) * L5 * site5)
            for bg in range(lo - lo % 256, hi, 256):            # seeds by global chunk of 256 trains: shard-count independent
                chunk = np.random.Generator(np.random.Philox(5000 + bg)).random(256 * L5 * site5)
                a, b_ = max(bg, lo), min(bg + 256, hi)
                packed[(a - lo) * L5 * site5:(b_ - lo) * L5 * site5] = chunk[(a - bg) * L5 * site5:(b_ - bg) * L5 * site5]
            return packed

        def c5_run(lo, hi, gather):
            """normalize! + pair marginals + compress! of trains [lo, hi) on this GPU; returns timings and results.
            The pair marginals are written to a DEVICE tensor (no pageable staging)."""
            nb_ = hi - lo
            A = T.PeriodicTensorTrain.empty([D5] * (L5 + 1), (Q5,), batch=nb_)
            packed_ = gen5(lo, hi)                              # generated once, uploaded twice (warm-up copy and timed copy)
            A.upload_all(packed_)
            T._ck(T.lib.ttb_twovar_count(A._hd, L5 - 1, C.byref(npair)))
            npr = int(npair.value)
            dev_out = torch.empty((nb_, npr, Q5 * Q5), dtype=torch.float64, device=dev)
            side = torch.cuda.Stream(device=dev)
            def pipeline(X, full_dev=None, host=None, gathered=False):
                """normalize! -> pair marginals (device tensor) -> compress!; the all-gather of the pair marginals and
                their copy into pinned host memory run on a side stream UNDER the compress! call"""
                ms = {}
                lz = T.normalize_(X); ms["normalize"] = X.last_timing()[0]
                T.twovar_marginals_block(X, 0, L5 - 1, L5 - 1, out=dev_out); ms["twovar_marginals"] = X.last_timing()[0]
                if host is not None:
                    work = dist.all_gather_into_tensor(full_dev, dev_out, async_op=True) if gathered else None
                    with torch.cuda.stream(side):
                        if work is not None:
                            work.wait()
                        if rank == 0 or not gathered:                  # the gathered result is read back once, on rank 0
                            host.copy_(full_dev, non_blocking=True)
                T.compress_(X, T.TruncThresh(1e-6)); ms["compress"] = X.last_timing()[0]
                side.synchronize()
                return lz, ms
            Wm = A.copy()
            pipeline(Wm)                                       # warm-up on the SAME handle (workspaces, scratch)
            Wm.reset(); Wm.upload_all(packed_)                 # fresh input in the warmed handle
            del packed_
            gathered = gather and world > 1 and B5 % world == 0
            host = pin_pairs((B5 if gathered else nb_, npr, Q5 * Q5))   # pinned destination exists before the clock starts
            full_dev = torch.empty((B5, npr, Q5 * Q5), dtype=torch.float64, device=dev) if gathered else dev_out
            if gather:
                barrier()
            else:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            lz, ms = pipeline(Wm, full_dev, host, gathered)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            if gathered and rank != 0:                               # (outside the clock: the in-run checks below want it everywhere)
                host.copy_(full_dev); torch.cuda.synchronize()
            bonds = np.array([Wm.bonds(b) for b in range(nb_)], dtype=np.int64)
            return {"A": A, "W": Wm, "lz": np.asarray(lz, dtype=np.float64), "ms": ms, "wall": wall, "pairs_host": host,
                    "bonds": bonds, "npr": npr}

        _pin = {}
        def pin_pairs(shape):
            key = tuple(shape)
            if key not in _pin:
                _pin.clear()
                _pin[key] = torch.empty(key, dtype=torch.float64, pin_memory=True)
            return _pin[key]

        lo5, hi5 = P.shard_range(B5, rank, world)
        nb = hi5 - lo5
        r5 = c5_run(lo5, hi5, gather=True)
        wall5 = max_over_ranks(r5["wall"])
        ms5 = r5["ms"]
        dev5 = max_over_ranks(sum(ms5.values()))
        ms5_max = {k: max_over_ranks(v) for k, v in ms5.items()}
        lz_all = P.allgather(np.ascontiguousarray(r5["lz"]).reshape(-1, 1)).reshape(-1) if (world > 1 and B5 % world == 0) else r5["lz"]
        bonds_all = P.allgather(r5["bonds"]) if (world > 1 and B5 % world == 0) else r5["bonds"]
        pairs_all = r5["pairs_host"].numpy()
        # HBM roofline of the environment pass behind normalize! (k_env_small_L reads every site tensor once:
        # algorithmic bytes = 8 * sum_t d_t d_{t+1} Q per train) and of the scaling pass (read + write)
        P5 = r5["A"].copy()
        P5.profile(True)
        T.normalize_(P5)
        rep5 = P5.profile_report()
        scratch_dev = torch.empty((nb, r5["npr"], Q5 * Q5), dtype=torch.float64, device=dev)
        T.twovar_marginals_block(P5, 0, L5 - 1, L5 - 1, out=scratch_dev)
        rep5b = P5.profile_report()
        T.compress_(P5, T.TruncThresh(1e-6))
        rep5c = P5.profile_report()
        P5.profile(False)
        del P5, scratch_dev
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs (burst copy figure)"
        except Exception:
            hbm_peak, hbm_src = 6545.9, "fallback: B200_PROFILING.md measured copy bandwidth"
        slab_bytes = float(nb) * L5 * site5 * 8
        roof5 = {}
        for kname, nbytes_alg in (("k_env_small_L", slab_bytes), ("k_scale_trains", 2 * slab_bytes)):
            if kname in rep5 and rep5[kname][1] > 0:
                gbs = nbytes_alg / (rep5[kname][1] * 1e-3) / 1e9
                roof5[kname] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                "ms": rep5[kname][1], "alg_bytes": nbytes_alg}
        # FP64-tensor roofline of the pair-marginal chain: per pair one stacked (Q d_1) x d x d update (the last
        # pair of every first site needs none) and the Q^2 Frobenius products; peak is filled in by rank 0 below
        if "k_twovar_mma" in rep5b and rep5b["k_twovar_mma"][1] > 0:
            npr = r5["npr"]
            fl5 = float(nb) * ((npr - (L5 - 1)) * 2.0 * Q5 * D5 * D5 * D5 + npr * 2.0 * Q5 * Q5 * D5 * D5)
            roof5["k_twovar_mma"] = {"bound": "tensor", "achieved": fl5 / (rep5b["k_twovar_mma"][1] * 1e-3) / 1e12, "unit": "TFLOP/s",
                                     "ms": rep5b["k_twovar_mma"][1], "alg_flops": fl5}
        compress_kernels = {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in sorted(rep5c.items(), key=lambda kv: -kv[1][1])
                            if k not in rep5b}
        assert np.all(np.isfinite(lz_all))
        batch_info = {"workload": "C5: 4096 PeriodicTensorTrains, 64 sites, bond 32, q=2, rand fill: normalize! + twovar_marginals "
                                  "(2016 pairs per train) + compress!(TruncThresh(1e-6)); trains sharded over ranks (strong scaling); "
                                  "all-gather (NCCL) of per-train log Z, bond dimensions and of the pair marginals between device "
                                  "buffers, then device -> pinned host",
                      "trains": B5, "trains_per_rank": nb, "device_ms": dev5, "device_ms_by_call_rank0": ms5,
                      "device_ms_by_call_max": ms5_max,
                      "trains_per_s": B5 / (dev5 * 1e-3), "e2e_wall_ms": wall5 * 1e3, "e2e_trains_per_s": B5 / wall5,
                      "compress_site_steps_per_s": B5 * 2 * L5 / (ms5_max["compress"] * 1e-3),
                      "logz_gathered": int(np.size(lz_all)), "pair_marginals_gathered_bytes": int(pairs_all.nbytes),
                      "roofline_rank0": roof5, "hbm_peak_source": hbm_src, "compress_kernels_rank0": compress_kernels,
                      "note": "device_ms = sum of the three calls' CUDA-event times, max over ranks; e2e_wall_ms = the three calls with the "
                              "all-gather of the pair marginals (device buffers, NCCL) and their copy into pinned host memory (rank 0) "
                              "running on a side stream under the compress! call"}
        c5_check = {"lz": lz_all, "bonds": bonds_all, "pairs": pairs_all.copy() if world > 1 else pairs_all}
        del r5
        if world > 1:
            one = alone_on_rank0(lambda: c5_run(0, B5, gather=False))
            if rank == 0:
                ms1 = sum(one["ms"].values())
                same = (np.array_equal(one["bonds"], c5_check["bonds"]) and
                        np.allclose(one["pairs_host"].numpy(), c5_check["pairs"], rtol=1e-12, atol=0) and
                        np.allclose(one["lz"], c5_check["lz"], rtol=1e-12, atol=0))
                pre_fail.append(None if same else "sharded C5 results differ from the one-GPU run")
                batch_info["one_gpu"] = {"device_ms": ms1, "device_ms_by_call": one["ms"], "e2e_wall_ms": one["wall"] * 1e3,
                                         "trains_per_s": B5 / (ms1 * 1e-3)}
                batch_info["speedup_vs_1gpu"] = {"device": ms1 / dev5, "e2e": one["wall"] / wall5,
                                                 "by_call": {k: one["ms"][k] / ms5_max[k] for k in ms5_max},
                                                 "results_identical_to_one_gpu_run": bool(same)}
            del one
        _pin.clear()

    # ---- third sharding axis: the pair marginals of ONE train -- the config-3 train itself (L=256, d=256,
    #      q=2: 32640 pairs) -- sharded over the ranks by first site, ragged all-gather between device buffers ----
    pairs_info = None
    if not args.no_extra:
        from ttb200 import parallel as P
        if rank == 0:
            A2 = base.copy()                                          # the seed-3 train of the headline leg
        else:                                                         # (every rank must hold the SAME train for this leg)
            _, t3 = make_open(3, L3, D3, Q3)
            A2 = T.TensorTrain(t3)
            del t3
        T.normalize_(A2)
        P.twovar_marginals_sharded(A2, as_dict=False)                 # warm-up
        barrier()
        t0 = time.perf_counter()
        full2 = P.twovar_marginals_sharded(A2, as_dict=False)
        wall2 = max_over_ranks(time.perf_counter() - t0)
        ms2 = max_over_ranks(A2.last_timing()[0])
        npairs2 = L3 * (L3 - 1) // 2
        assert full2.shape == (npairs2, Q3 * Q3) and np.allclose(full2.sum(axis=1), 1.0, rtol=0, atol=1e-9)
        pairs_info = {"workload": "pair marginals of ONE TensorTrain (the config-3 train, L=256 q=2 d=256, normalized), all 32640 pairs, "
                                  "first sites sharded over the ranks (parallel.pair_ranges: equal pair counts), ragged NCCL all-gather "
                                  "of the blocks from device buffers into pinned host memory",
                      "pairs": npairs2, "device_ms": ms2, "wall_ms": wall2 * 1e3, "pairs_per_s": npairs2 / (ms2 * 1e-3),
                      "e2e_pairs_per_s": npairs2 / wall2,
                      "note": "one CTA per first site walks its chain of later sites, so the longest chain (first site 0: L-1 steps) "
                              "bounds the call on every rank that holds an early first site; the environments are recomputed per rank"}
        if world > 1:
            def one2():
                T.twovar_marginals_block(A2, 0, L3 - 1)
                t0_ = time.perf_counter()
                o = T.twovar_marginals_block(A2, 0, L3 - 1)
                return o[0], A2.last_timing()[0], time.perf_counter() - t0_
            one = alone_on_rank0(one2)
            if rank == 0:
                same3 = bool(np.allclose(one[0], full2, rtol=1e-12, atol=0))
                pre_fail.append(None if same3 else "sharded pair marginals of one train differ from the one-GPU run")
                pairs_info["results_identical_to_one_gpu_run"] = same3
                pairs_info["one_gpu"] = {"device_ms": one[1], "wall_ms": one[2] * 1e3}
                pairs_info["speedup_vs_1gpu"] = {"device": one[1] / ms2, "e2e": one[2] / wall2}
        del A2

    if rank == 0:
        # ---- roofline: per-kernel CUDA-event times of one more compress! (serial timeline) ----
        prof = base.copy()
        prof.profile(True)
        T.compress_(prof, tr)
        rep = prof.profile_report()
        prof.profile(False)
        peak = measure_fp64_peak(torch, dev)
        fl, _ = qr_sweep_flops(bonds, Q3)
        alias = {"k_qr_panel_flag": "k_qr_panel", "k_qr_update_tma": "k_qr_update", "k_apply_q_tma": "k_apply_q",
                 "k_absorb_mma": "k_absorb"}
        tot_ms = sum(v[1] for v in rep.values())
        dom = max(rep, key=lambda k: rep[k][1])
        n_l, ms = rep[dom]
        # algorithmic FLOPs of the dominant class over the step (right sweep; the left sweep of the
        # rand fill works on <=16-row matrices and is negligible) -- DESIGN.md states the formulas
        n_alg, alg = fl.get(alias.get(dom, dom), [n_l, 0.0])
        per_launch = alg / max(n_alg, 1)
        achieved = alg / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        # DRAM traffic per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of the committed
        # `ncu --set full` capture, read from profiles/ncu_traffic.json (profiles/make_ncu_traffic.py writes it from the
        # per-round summaries); null when the kernel has no capture
        try:
            ncu_tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            ncu_tab = {}
        ncu_traffic = {k: int(v["dram_bytes_per_launch"]) for k, v in ncu_tab.items()}
        ncu_src = {k: v.get("source") for k, v in ncu_tab.items()}
        line["roofline"] = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak if peak else None, "traffic": ncu_traffic.get(dom),
                            "traffic_source": ncu_src.get(dom),
                            "peak_source": "measured in this run: torch.matmul float64 4096^3 (cuBLAS DGEMM, best of 8); "
                                           "MEASURED_PEAKS.json has no FP64 entry",
                            "share_of_step": ms / tot_ms if tot_ms else None,
                            "flops_per_launch": per_launch, "launches": n_l, "working_launches": n_alg,
                            "ms_per_launch": ms / n_l,
                            "note": "k_qr_stream = the whole Householder QR of one site (512 x 256) as one 8-CTA cluster: 256 "
                                    "chained columns, dependency-latency bound by construction (profiles/r1_micro_panel.txt, "
                                    "r1_ncu_panel_warpstates.txt); achieved = algorithmic QR FLOPs 2k^2(p-k/3) / CUDA-event "
                                    "time of all launches of the class; traffic = dram bytes of one launch under ncu --set full "
                                    "(the 1.05 MB site read once, writes absorbed by L2)",
                            "kernels": {k: {"launches": v[0], "ms": v[1],
                                            "alg_gflop": fl.get(alias.get(k, k), [0, 0.0])[1] / 1e9}
                                        for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])},
                            "whole_step": {"alg_gflop": sum(v[1] for v in fl.values()) / 1e9,
                                           "tflops": sum(v[1] for v in fl.values()) / (dev_ms / K * 1e-3) / 1e12}}
        extra = {}
        parity_fail = any(m is not None for m in pre_fail)
        if parity_fail:
            line["parity_failures"] = [m for m in pre_fail if m is not None]
        if sample_info is not None:
            sample_info["tflops_minimal_algorithm"] = sample_info["draws_per_s"] * 8.65e6 / 1e12
            sample_info["frac_of_fp64_peak"] = sample_info["tflops_minimal_algorithm"] / (peak * world) if peak else None
            sample_info["roofline"] = {"bound": "tensor", "kernel": "k_sample_tile<128>", "unit": "TFLOP/s",
                                       "achieved": sample_info["tflops_minimal_algorithm"] / world, "peak": peak,
                                       "frac": sample_info["frac_of_fp64_peak"],
                                       "ncu": "tensor pipe 64.9% active, DRAM traffic 34.1 MB per launch = the packed "
                                              "train read once (profiles/r1_ncu_full_sample_tile.csv)"}
            extra["sample"] = sample_info
        if batch_info is not None:
            tw = batch_info.get("roofline_rank0", {}).get("k_twovar_mma")
            if tw is not None and peak:
                tw["peak"] = peak
                tw["frac"] = tw["achieved"] / peak
            extra["batch"] = batch_info
        if pairs_info is not None:
            extra["pairs_of_one_train"] = pairs_info
        if not args.no_extra:
            # ---- config 2 (BASELINE.json configs[1]): single TT L=128 q=2 d=64: orthogonalize_right! + compress!(TruncBond(32))
            #      + normalize! + all single-site marginals, device time of the four calls ----
            try:
                L2, D2, Q2 = 128, 64, 2
                _, t2 = make_open(2, L2, D2, Q2)
                def c2_pipeline(A):
                    ms = {}
                    T.orthogonalize_right_(A, T.TruncThresh(0.0)); ms["orthogonalize_right"] = A.last_timing()[0]
                    T.compress_(A, T.TruncBond(32), is_orthogonal="right"); ms["compress_TruncBond32"] = A.last_timing()[0]
                    lz = T.normalize_(A); ms["normalize"] = A.last_timing()[0]
                    m = T.marginals(A); ms["marginals"] = A.last_timing()[0]
                    return ms, lz, m
                A2 = T.TensorTrain(t2)
                c2_pipeline(A2.copy())
                w0_ = time.perf_counter()
                ms2, lz2, m2 = c2_pipeline(A2)
                wall2_ = time.perf_counter() - w0_
                c2 = {"workload": "C2: single TensorTrain L=128 q=2 d=64 rand_tt: orthogonalize_right!(TruncThresh(0.0)) + "
                                  "compress!(TruncBond(32), is_orthogonal=:right) + normalize! + marginals (254 site-steps)",
                      "device_ms": sum(ms2.values()), "device_ms_by_call": ms2, "wall_ms": wall2_ * 1e3,
                      "site_steps_per_s": 2 * (L2 - 1) / ((ms2["orthogonalize_right"] + ms2["compress_TruncBond32"]) * 1e-3),
                      "bonds_after_max": int(max(T.bond_dims(A2)))}
                extra["config2"] = c2
            except Exception as ex:
                extra["config2_error"] = repr(ex)
        if world == 1 and not args.no_extra:
            # in a clean process (this one holds a CUDA context, pinned buffers and sampler threads, which was
            # measured to slow the host BLAS several-fold): exactly what `--impl reference` times
            refl = None
            try:
                cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                                    capture_output=True, text=True, timeout=900)
                refl = json.loads([ln for ln in cp.stdout.splitlines() if ln.startswith("{")][-1])
                line["cpu_baseline"] = dict(refl["cpu_baseline"], measured="subprocess: bench.py --impl reference --steps 1 --warmup 1")
            except Exception as ex:
                v, secs = cpu_compress_steps_per_s(L3, D3, Q3, 3)
                line["cpu_baseline"] = {"value": v, "unit": "site-steps/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"the full C3 compress! once in-process ({secs:.1f} s; subprocess failed: {ex!r}): "
                                                  "NumPy/SciPy-OpenBLAS gesdd restatement of the reference, not Julia/MKL"}
                refl = {"config": {"bonds_after": LAST_CPU_BONDS}}
            # parity of the TIMED run against the CPU arm, outside the timed region: retained bond dimensions of the
            # last timed compress! (same seed-3 input) must be identical to the oracle's
            cpu_bonds = (refl or {}).get("config", {}).get("bonds_after")
            if cpu_bonds is not None:
                same = [int(x) for x in bonds_after] == [int(x) for x in cpu_bonds]
                line["parity_in_run"] = {"bonds_identical_to_cpu_arm": bool(same), "bonds_after": [int(x) for x in bonds_after],
                                         "what": "retained bond dimensions of the last timed compress! vs the oracle's on the same input "
                                                 "(the CPU arm's own run)"}
                parity_fail = parity_fail or not same
            try:
                try:   # same BLAS thread count as the compress baseline picked
                    from threadpoolctl import threadpool_limits
                    _lim = threadpool_limits(limits=int(line["cpu_baseline"].get("cores", os.cpu_count())))
                except Exception:
                    _lim = None
                extra["sample"]["cpu_draws_per_s"] = cpu_sample_draws_per_s(L4, D4, Q4, 4, 20)
                extra["sample"]["cpu_blas_threads"] = int(line["cpu_baseline"].get("cores", 0))
                # ---- C2 on the CPU (oracle), same pipeline ----
                if "config2" in extra:
                    from oracle import tt_oracle as O
                    _, t2 = make_open(2, 128, 64, 2)
                    Ao = O.TensorTrain(t2)
                    c0_ = time.perf_counter()
                    O.orthogonalize_right(Ao, O.TruncThresh(0.0)); O.compress(Ao, O.TruncBond(32), is_orthogonal="right")
                    c1_ = time.perf_counter()
                    lzo = O.normalize(Ao); mo = O.marginals(Ao)
                    c2_ = time.perf_counter()
                    extra["config2"]["cpu_ms"] = (c2_ - c0_) * 1e3
                    extra["config2"]["cpu_site_steps_per_s"] = 254 / (c1_ - c0_)
                    extra["config2"]["matches_cpu"] = bool(T.bond_dims(A2) == O.bond_dims(Ao) and abs(lz2 - lzo) <= 1e-8 * max(1.0, abs(lzo))
                                                          and all(np.allclose(a, b, rtol=1e-8, atol=1e-12) for a, b in zip(m2, mo)))
                    parity_fail = parity_fail or not extra["config2"]["matches_cpu"]
                # ---- C5: fixed global train ids against the oracle (log Z, every pair marginal, retained bonds) ----
                if batch_info is not None:
                    from oracle import tt_oracle as O
                    ids = [0, 1, 777, 2047, 2048, 4095]
                    ok_ids = []
                    for g in ids:
                        raw = gen5(g, g + 1).reshape(L5, D5 * D5 * Q5)
                        Ao = O.PeriodicTensorTrain([raw[t].reshape((D5, D5, Q5), order="F").copy() for t in range(L5)])
                        lzo = O.normalize(Ao)
                        po = O.twovar_marginals(Ao)
                        O.compress(Ao, O.TruncThresh(1e-6))
                        okg = abs(c5_check["lz"][g] - lzo) <= 1e-10 * abs(lzo)
                        i = 0
                        for t in range(L5 - 1):
                            for u in range(t + 1, L5):
                                okg = okg and np.allclose(c5_check["pairs"][g, i], np.reshape(po[(t, u)], -1, order="F"), rtol=1e-9, atol=1e-13)
                                i += 1
                        okg = okg and [int(x) for x in c5_check["bonds"][g][:L5]] == [int(x) for x in O.bond_dims(Ao)]
                        ok_ids.append(bool(okg))
                    batch_info["oracle_check"] = {"global_train_ids": ids, "ok": ok_ids,
                                                  "what": "log Z (1e-10), all 2016 pair marginals (1e-9), retained bonds (identical) vs the oracle"}
                    parity_fail = parity_fail or not all(ok_ids)
                # ---- randn fill of config 3 at FULL length (both sweeps full rank: every left-sweep step is a 256 x 256 SVD) ----
                br, tsr = make_open(3, L3, D3, Q3, "randn")
                Ar = T.TensorTrain(tsr)
                del tsr
                T.compress_(Ar.copy(), tr)
                Pr = Ar.copy()
                T.compress_(Ar, tr)
                msr, _ = Ar.last_timing()
                Pr.profile(True); T.compress_(Pr, tr); repr_ = Pr.profile_report(); Pr.profile(False)
                bonds_r = T.bond_dims(Ar)
                del Pr
                nsvd = L3 - 1
                f_svd = 6.0 * 512 * 256 ** 2 + 20.0 * 256 ** 3          # GVL R-SVD, thin U, S, V of the 512 x 256 fold
                jac = {k: v for k, v in repr_.items() if "jacobi" in k}
                jac_ms = sum(v[1] for v in jac.values())
                rn = {"workload": f"C3 shape, randn_tt fill (nothing to truncate: every left-sweep step is a full 256 x 256 SVD), L={L3}, "
                                  "compress!(TruncThresh(1e-10)) = 510 site-steps",
                      "site_steps_per_s": nsteps_unit / (msr * 1e-3), "ms": msr, "bonds_after_max": int(max(bonds_r)),
                      "kernels": {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in sorted(repr_.items(), key=lambda kv: -kv[1][1])},
                      "roofline": {"bound": "tensor", "kernel": "k_jacobi_* (QR + one-sided Jacobi SVD of the left sweep)", "unit": "TFLOP/s",
                                   "achieved": nsvd * f_svd / (jac_ms * 1e-3) / 1e12 if jac_ms else None, "peak": peak,
                                   "frac": nsvd * f_svd / (jac_ms * 1e-3) / 1e12 / peak if jac_ms and peak else None,
                                   "alg_flops_per_step": f_svd, "ms": jac_ms,
                                   "note": "achieved = 255 x F_svd (6 n m^2 + 20 m^3 = 536.9 MFLOP, the LAPACK R-SVD count SURVEY 8d prescribes; "
                                           "the absorb's 67.1 MFLOP is counted with its own kernel) / CUDA-event time of the Jacobi kernels"}}
                try:
                    cpr = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--fill", "randn", "--steps", "1",
                                          "--warmup", "0"], capture_output=True, text=True, timeout=900)
                    rl = json.loads([ln for ln in cpr.stdout.splitlines() if ln.startswith("{")][-1])
                    rn["cpu_site_steps_per_s"] = rl["value"]
                    rn["cpu_sample"] = rl["cpu_baseline"]["sample"]
                    rn["cpu_blas_threads"] = rl["cpu_baseline"]["cores"]
                    if rl["config"].get("bonds_after") is not None:
                        rn["bonds_identical_to_cpu_arm"] = [int(x) for x in bonds_r] == [int(x) for x in rl["config"]["bonds_after"]]
                        parity_fail = parity_fail or not rn["bonds_identical_to_cpu_arm"]
                except Exception as ex:
                    rn["cpu_error"] = repr(ex)
                extra["randn_fill"] = rn
            except Exception as ex:  # the headline line must still be printed
                extra["error"] = repr(ex)
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
        sys.stdout.flush()
        if parity_fail:
            raise SystemExit("bench.py: a parity check of this run FAILED (see parity_in_run / matches_cpu / oracle_check in the line)")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
