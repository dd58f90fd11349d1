#!/usr/bin/env python
"""bench.py — the headline benchmark of the hot path (BASELINE.json `metric`).

Workload (config.workload): single SpGEMM C = A·A on a synthetic directed R-MAT graph, scale 20,
edge factor 16 (BASELINE.json configs[1]); rmclInit semantics (self loops, values 1/rowcount).
A "step" is one whole pass of the path over that input: flops analysis → binning → symbolic →
row offsets → numeric (sorted columns), result left in HBM.

  python bench.py [--gpus N] [--steps K] [--warmup W]      one JSON line from rank 0
  python bench.py --impl reference ...                       the reference's OpenMP CPU path

* value       2·P·K / t, P = intermediate products (counted by the flops-analysis kernel), t =
              CUDA-event time of K steps on the library's stream, max over ranks; operands
              resident in HBM.  N > 1: the rows of A are cut into N contiguous blocks of equal
              cost (arrayEqualPartition64 on products + a per-row charge), rank r computes
              block r against the full B; no collective on the data path; total work fixed =>
              "scaling": "strong".
* e2e         the same product through the reference-facing host-buffer call: malloc'd int
              CSR in, malloc'd int CSR out, H2D and D2H copies inside the timed region.
              nnz(C) = 9.7e9 exceeds the reference's `int` CSR, so the result arrives as
              consecutive row blocks whose products (an upper bound of their nnz) fit an int
              (SURVEY.md §7): b200_spgemm_csr_stream at N = 1 (the library cuts the blocks and
              overlaps the download of one with the computation of the next), the
              toGpuCSR -> gpuSpMMWrapper(row block) -> toCpuCSR(row block) sequence per rank
              at N > 1.  The caller hands every block back with b200_host_free.
* roofline    the kernel with the largest share of the step: algorithmic bytes of the rows it
              processed ÷ its CUDA-event duration (b200_stats.ms_*_bin), against the measured
              HBM copy bandwidth of MEASURED_PEAKS.json.
* cpu_baseline / --impl reference
              the UNMODIFIED reference flops_omp_CSR_SpMM (oracle/_ref/libref.so, built by
              oracle/Makefile from /root/reference) — or the C restatement oracle/liboracle.so
              when that file is absent — on all host cores, on a bounded row sample of the same
              product (a hashed 1/q of the rows of A against the full B; the sample's product
              count fits the reference's `int` CSR).

Nothing on the measured GPU path touches oracle/.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time

# Host threads.  torchrun exports OMP_NUM_THREADS=1; the CPU legs (the reference's OpenMP SpGEMM)
# and the library's host-side staging copies are OpenMP code, so give every rank its share of
# the cores — all of them for the reference arm, which runs on rank 0 alone — before any
# OpenMP runtime is loaded.
_world = int(os.environ.get("WORLD_SIZE", "1"))
_is_ref = "reference" in sys.argv
try:
    _cores = sorted(os.sched_getaffinity(0))
except AttributeError:
    _cores = list(range(os.cpu_count() or 1))
_share = max(1, len(_cores) // (1 if _is_ref else _world))
os.environ["OMP_NUM_THREADS"] = str(_share)
# BASELINE.md §3 / SURVEY.md §8d timing protocol: OMP_PROC_BIND=close.  With several ranks on one
# host every rank needs its OWN places: `close` alone binds every process's threads — the
# thread that launches the kernels included — to the same first cores (measured, round 2, N = 8:
# 68 ms per step instead of 34, end-to-end 18 s instead of 6).
os.environ.setdefault("OMP_PROC_BIND", "close")
if not _is_ref and _world > 1 and "OMP_PLACES" not in os.environ:
    _lr = int(os.environ.get("LOCAL_RANK", "0"))
    _mine = _cores[_lr * _share:(_lr + 1) * _share] or _cores
    os.environ["OMP_PLACES"] = ",".join("{%d}" % c for c in _mine)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spgemm_gflops"
UNIT = "GFLOP/s"
L2_BYTES = 126 * 1024 * 1024

# numeric / symbolic bin names (sparse_matrix_with_flops_b200/csrc/spgemm.cu)
SYM_KERNELS = {1: "k_sym_warp<256>", 2: "k_sym_warp<1024>", 3: "k_sym_warp<4096>",
               4: "k_sym_warp<16384>", 5: "k_sym_bitmap",
               6: "k_esc<256,8> (expand/sort/compress, SpGEMM)", 7: "k_esc<512,16> (expand/sort/compress, SpGEMM)"}
NUM_KERNELS = {1: "k_num_warp<64>", 6: "k_num_warp<128>", 2: "k_num_warp<256>", 3: "k_num_warp<1024>",
               4: "k_num_warp<2048>", 5: "k_num_bitmap", 7: "k_esc<256,8> (rMCL) / k_esc_gather",
               8: "k_esc<512,16> (rMCL) / k_esc_gather", 9: "k_esc_gather (rows of k_num_warp_fused)"}
FUSED_BIN = 9   # numeric bin of the rows k_num_warp_fused<128> finished in the symbolic phase


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="rmat20",
                    help="A*A: rmat<scale> (directed, edge factor 16) | stencil<g> (g^3 27-point) | "
                         "planted<n>c<clusters>;  rMCL loop: rmcl-rmat<scale>[e<edge factor>] "
                         "(symmetrised) | rmcl-planted<n>c<clusters>")
    ap.add_argument("--rmcl-iters", type=int, default=5,
                    help="iterations of one rMCL loop (the reference's default maxIters, process_args.h:28)")
    ap.add_argument("--rmcl-leg", default="rmcl-rmat20",
                    help="rMCL workload measured next to an A*A headline and reported under \"rmcl\" "
                         "(BASELINE metric: SpGEMM GFLOP/s + rMCL iter/s); 'none' skips it")
    ap.add_argument("--no-e2e", action="store_true", help="development: skip the host-buffer leg")
    ap.add_argument("--no-cpu", action="store_true", help="development: skip the CPU baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-api", default="auto", choices=["auto", "streamed", "sequential"],
                    help="e2e leg: b200_spgemm_csr_stream (download of block b overlaps the product of block "
                         "b+1) or upload / product / download per row block; auto = streamed on one GPU, "
                         "sequential with several ranks on one host")
    ap.add_argument("--no-pin", action="store_true",
                    help="e2e leg: plain malloc result blocks (the library's default) instead of the "
                         "page-locked block cache (b200_host_cache_pin)")
    ap.add_argument("--row-charge", type=int, default=32768,
                    help="products-equivalent fixed cost of a heavy row in the N>1 row partition")
    return ap.parse_args()


def make_workload(smf, name):
    if name.startswith("rmat"):
        scale = int(name[4:])
        A = smf.synth_rmat(scale, 16, 12345, False)
        desc = "C=A*A, R-MAT scale %d edge factor 16 directed (a,b,c,d=.57,.19,.19,.05), seed 12345" % scale
    elif name.startswith("stencil"):
        g = int(name[7:])
        A = smf.synth_stencil27(g, g, g)
        desc = "C=A*A, 27-point stencil on a %d^3 grid" % g
    elif name.startswith("planted"):
        n, k = (int(x) for x in name[7:].split("c"))
        A = smf.synth_planted(n, k, 16, 2, 12345)
        desc = "C=A*A, planted partition %d vertices %d blocks (16 intra + 2 inter draws per vertex)" % (n, k)
    else:
        raise SystemExit("unknown workload " + name)
    return A, desc


def make_rmcl_workload(smf, name):
    """rMCL inputs: Mgt = Mt0 = rmclInit of a symmetrised, de-duplicated graph."""
    body = name[len("rmcl-"):]
    if body.startswith("rmat"):
        parts = body[4:].split("e")
        scale, ef = int(parts[0]), (int(parts[1]) if len(parts) > 1 else 16)
        A = smf.synth_rmat(scale, ef, 12345, True)
        desc = "rMCL loop on R-MAT scale %d edge factor %d symmetrised, seed 12345" % (scale, ef)
        small = "rmcl-rmat%de%d" % (min(scale, 15), ef)
    elif body.startswith("planted"):
        n, k = (int(x) for x in body[7:].split("c"))
        A = smf.synth_planted(n, k, 16, 2, 12345)
        desc = "rMCL loop on planted partition %d vertices %d blocks" % (n, k)
        ns = min(n, 100000)
        small = "rmcl-planted%dc%d" % (ns, max(1, k * ns // n))
    else:
        raise SystemExit("unknown rMCL workload " + name)
    return A, desc, small


def rmcl_iter_bytes(n, nnzG, products, nnz_new):
    """ALGORITHMIC bytes of one fused rMCL iteration: A = Mgt rows, gathered
    B = Mt rows (12 B per product + the row-pointer pair per A entry), C = the PRUNED new Mt."""
    return 12 * nnzG + 4 * (n + 1) + 12 * products + 8 * nnzG + 12 * nnz_new + 4 * (n + 1)


def host_flops_prefix(A):
    """Per-row intermediate products of A·A as an exclusive prefix (host numpy; used by the
    reference arm and to cut row blocks for the host-buffer API)."""
    rowlen = np.diff(A.rowPtr).astype(np.int64)
    per_entry = rowlen[A.colInd]
    cs = np.concatenate([[0], np.cumsum(per_entry)])
    return cs[A.rowPtr.astype(np.int64)]


# ---- CPU baseline (checker libraries; never on the GPU path) --------------------------------

def row_sample(A, q):
    """A pseudo-random 1/q of the rows of A (Fibonacci hash of the row id; a plain stride would
    over-sample R-MAT's hubs, whose ids end in zero bits) as a CSR over the same columns."""
    ids = np.arange(A.rows, dtype=np.uint64)
    rows = np.nonzero(((ids * np.uint64(2654435761)) % np.uint64(1 << 32)) < np.uint64((1 << 32) // q))[0].astype(np.int64)
    starts, ends = A.rowPtr[rows].astype(np.int64), A.rowPtr[rows + 1].astype(np.int64)
    lens = ends - starts
    I = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    idx = np.repeat(starts - I[:-1], lens) + np.arange(int(I[-1]), dtype=np.int64)
    return I, np.ascontiguousarray(A.colInd[idx]), np.ascontiguousarray(A.values[idx]), len(rows)


class CpuArm:
    """flops_omp_CSR_SpMM of the reference (kind 'reference') or the oracle port (kind 'port')."""

    def __init__(self, A, target_products=1.5e9):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        self.ol = ol
        self.A = A
        prefix = host_flops_prefix(A)
        P = int(prefix[-1])
        rowlen = np.diff(A.rowPtr).astype(np.int64)
        self.q = max(1, int(np.ceil(P / target_products)))
        while True:
            # nnz(C) <= products must fit the reference's `int` CSR (nlibs/CSR.h:38)
            self.I, self.J, self.V, self.m = row_sample(A, self.q)
            self.products = int(rowlen[self.J].sum())
            if self.products <= 2_000_000_000:
                break
            self.q += 1
        self.kind = "reference" if os.path.exists(ol.REF_SO) else "port"
        self.cores = os.cpu_count()
        self.sample = ("hashed 1/%d row sample of A (%d rows, %d of %d products) x full B" %
                       (self.q, self.m, self.products, P))

    def step_ms(self):
        A, ol = self.A, self.ol
        ip, dp = ol.ip, ol.dp
        _i, _d = ol._i, ol._d
        if self.kind == "reference":
            r = ol.ref()
            nnzc = C.c_longlong(0)
            self.cores = r.ref_num_threads()
            ms = r.ref_spgemm_timed(3, _i(self.I), _i(self.J), _d(self.V), int(self.I[-1]),
                                    _i(A.rowPtr), _i(A.colInd), _d(A.values), A.nnz, self.m,
                                    A.cols, A.cols, 512, 1, C.byref(nnzc))
            return float(ms)
        o = ol.oracle()
        IC, JC, Cv, n = ip(), ip(), dp(), C.c_int()
        t0 = time.perf_counter()
        rc = o.oracle_spgemm(_i(self.I), _i(self.J), _d(self.V), _i(A.rowPtr), _i(A.colInd),
                             _d(A.values), self.m, A.cols, C.byref(IC), C.byref(JC), C.byref(Cv),
                             C.byref(n))
        ms = (time.perf_counter() - t0) * 1e3
        assert rc == 0, rc
        for p in (IC, JC, Cv):
            o.oracle_free(C.cast(p, C.c_void_p))
        return ms

    def gflops(self, ms):
        return 2.0 * self.products / (ms * 1e-3) / 1e9


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import sparse_matrix_with_flops_b200 as smf  # input generator only (host code)
    A, desc = make_workload(smf, args.workload)
    arm = CpuArm(A)
    for _ in range(args.warmup):
        arm.step_ms()
    times = [arm.step_ms() for _ in range(args.steps)]
    ms = sum(times) / len(times)
    val = arm.gflops(ms)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "timing": "omp_get_wtime around flops_omp_CSR_SpMM, thread scratch allocated outside"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                         "sample": arm.sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks ------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock and throttle reasons during the timed region, through NVML in this process (a
    background thread; an `nvidia-smi -lms` child proved to stall kernel launches by ~100 ms per
    query on this driver, so it is not used)."""
    HW_SLOWDOWN, SW_THERMAL, HW_THERMAL, SW_POWER_CAP = 0x8, 0x20, 0x40, 0x4

    def __init__(self, gpu_index, period_s=0.1):
        import threading
        self.samples, self.reasons, self.err = [], set(), None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # no NVML: report it, do not fail the bench
            self.err = repr(e)
            return
        self.period = period_s
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, pw))
                for bit, nm in ((self.HW_SLOWDOWN, "hw_slowdown"), (self.SW_THERMAL, "sw_thermal_slowdown"),
                                (self.HW_THERMAL, "hw_thermal_slowdown"), (self.SW_POWER_CAP, "sw_power_cap")):
                    if r & bit:
                        self.reasons.add(nm)
            except Exception as e:
                self.err = repr(e)
                return
            self._stop.wait(self.period)

    def stop(self):
        if self.err and not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + self.err]}
        self._stop.set()
        self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(x[0] for x in self.samples), "sm_max_mhz": self.max_sm,
                "power_w_max": max(x[1] for x in self.samples), "samples": len(self.samples),
                "reasons": sorted(self.reasons)}



# ---- rMCL loop (BASELINE metric, second half: rMCL iterations / second) -----------------------

RMCL_METRIC = "rmcl_iters_per_s"
RMCL_UNIT = "iter/s"


class quiet_stdout:
    """The reference's loop prints a line per iteration (nlibs/qrmcl.cc:66-70) on the C stdout;
    the bench line must stay the only thing on ours."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:  # noqa: BLE001
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def rmcl_cpu_baseline(smf, small_name, iters, full_products_per_iter):
    """The reference's own loop (mtRmclIter with static_omp_CSR_RMCL_OneStep, nlibs/qrmcl.cc:8-84
    through oracle/_ref; the checker's port when the reference library is absent) on a smaller
    graph of the same family — the full one does not fit the reference's `int` CSR — timed on
    the host cores.  Its products / second, divided by the full workload's products per
    iteration, is the CPU's equivalent iterations / second on the full workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    As, desc_s, _ = make_rmcl_workload(smf, small_name)
    M = ol.from_csr(As)
    # products of the sample loop, from the checker's flops analysis of every iterate
    kind = "reference" if os.path.exists(ol.REF_SO) else "port"
    if kind == "reference":
        cores = ol.ref().ref_num_threads()
        with quiet_stdout():
            ol.r_rmcl_iter(M, M, 1)                   # warm-up (thread scratch, page faults)
            ms = ol.r_rmcl_iter(M, M, iters).ms
    else:
        cores = 1
        t0 = time.perf_counter()
        ol.o_rmcl_iter(M, M, iters)
        ms = (time.perf_counter() - t0) * 1e3
    T, prods = M, 0
    for _ in range(iters):
        prods += int(ol.o_flops_prefix(M, T)[-1])
        T = ol.o_rmcl_onestep(M, T)
    pps = prods / (ms * 1e-3)
    return {"value": pps / full_products_per_iter, "unit": RMCL_UNIT, "cores": cores, "kind": kind,
            "sample": "%s: %d iterations, %d products in %.1f ms (%.3g products/s = %.3f iter/s there); value = "
                      "that rate / the full workload's %.4g products per iteration" % (
                          desc_s, iters, prods, ms, pps, iters / (ms * 1e-3), full_products_per_iter),
            "sample_iters_per_s": iters / (ms * 1e-3), "products_per_s": pps, "ms": ms}


def rmcl_measure(args, env, name, steps, warmup, want_cpu, want_e2e):
    """K rMCL loops of --rmcl-iters iterations each (a step = one loop from Mt0 = Mgt) through
    b200_rmcl_iter_sharded: at N > 1 every iteration cuts the rows into N flops-balanced
    blocks, all-gathers the pruned blocks over NCCL and max-reduces chaos."""
    import torch
    import torch.distributed as dist
    smf, lib, stream, barrier, rank, world = (env[k] for k in ("smf", "lib", "stream", "barrier", "rank", "world"))
    A, desc, small = make_rmcl_workload(smf, name)
    iters = args.rmcl_iters
    if world > 1 and not env.get("comm"):
        uid = [smf.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        with quiet_stdout():
            smf.comm_init(rank, world, uid[0])
        env["comm"] = True
    dG = A.toGpuCSR()

    def loop(dT):
        return smf.gpuRmclIterSharded(iters, dG, dT, want_counts=True)

    warm_iter_ms = []
    for _ in range(warmup):
        dT = A.toGpuCSR()
        warm_iter_ms.append([round(float(x), 1) for x in loop(dT)[2]])
        dT.deviceDispose()
    dTs = [A.toGpuCSR() for _ in range(steps)]     # Mt0 of every timed loop: resident before the clock starts
    barrier()
    sampler = ClockSampler(env["local"]) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    launches = 0
    loops_ms, loops_iter_ms = [], []
    for dT in dTs:
        done, hist, ms_it, counts = loop(dT)
        launches += int(counts[:, 4].sum())
        loops_ms.append(round(float(ms_it.sum()), 3))
        loops_iter_ms.append([round(float(x), 1) for x in ms_it])
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    nnz_final = dTs[-1].info()[2]
    for dT in dTs:
        dT.deviceDispose()
    t = torch.tensor([ms_total, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_total, launches = float(tmax[0]), int(t[1])
    ms_loop = ms_total / steps
    value = done / (ms_loop * 1e-3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    bytes_it = [rmcl_iter_bytes(A.rows, A.nnz, int(counts[k, 0]), int(counts[k, 1])) for k in range(done)]
    achieved = sum(bytes_it) / (ms_loop * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "one rMCL loop (every kernel of its %d fused iterations)" % done,
                "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                "traffic": None, "peak_source": peak_src + (" x %d GPUs" % world if world > 1 else ""),
                "loop_algorithmic_bytes": int(sum(bytes_it)),
                "loops_ms_rank0": loops_ms,   # sum of the library's per-iteration CUDA-event times, per timed loop
                "loops_iteration_ms_rank0": loops_iter_ms,
                "warmup_loops_iteration_ms_rank0": warm_iter_ms,
                "per_iteration": {"ms": [round(float(x), 3) for x in ms_it],
                                  "products": [int(x) for x in counts[:, 0]],
                                  "nnz_new_Mt": [int(x) for x in counts[:, 1]],
                                  "nnz_unpruned": [int(x) for x in counts[:, 2]],
                                  "row_tiles": [int(x) for x in counts[:, 3]],
                                  "chaos": [float(x) for x in hist]}}
    # ---- e2e: host CSR in, host CSR out, every copy inside the clock
    e2e = None
    if want_e2e:
        def one():
            if world == 1:
                Mt, it, _ = smf.gpuRmclIter(iters, A, A)
                return Mt.nnz
            g, tt = A.toGpuCSR(), A.toGpuCSR()
            smf.gpuRmclIterSharded(iters, g, tt)
            Mt = tt.toCpuCSR()
            g.deviceDispose(); tt.deviceDispose()
            return Mt.nnz
        one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            nz = one()
        barrier()
        sec = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt[0])
        per = 4 * (A.rows + 1) + 12 * A.nnz
        e2e = {"value": iters / sec, "unit": RMCL_UNIT, "h2d_bytes_per_step": 2 * per * world,
               "d2h_bytes_per_step": (4 * (A.rows + 1) + 12 * nz) * world, "ms_per_step": sec * 1e3,
               "steps": args.e2e_steps, "warmup": 1,
               "api": ("b200_rmcl_iter (host int CSR of Mgt and Mt in, malloc'd host int CSR of the final Mt out)"
                       if world == 1 else
                       "per rank: b200_csr_upload x2 + b200_rmcl_iter_sharded + b200_csr_download") +
                      ", wall clock incl. H2D + D2H"}
    cpu = None
    if want_cpu and rank == 0 and world == 1:
        cpu = rmcl_cpu_baseline(smf, small, iters, float(np.mean(counts[:, 0])))
    dG.deviceDispose()
    return {"metric": RMCL_METRIC, "value": value, "unit": RMCL_UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_loop, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "rows": A.rows, "nnz_Mgt": A.nnz, "iterations_per_loop": done,
                       "nnz_final": int(nnz_final),
                       "partition": "flops-balanced contiguous row blocks of Mgt, recomputed every iteration (arrayEqualPartition64)",
                       "collectives": ("none (1 rank)" if world == 1 else
                                       "per iteration, in the library's own NCCL communicator (%d ranks): ncclAllGather of "
                                       "{block nnz, status, chaos} + grouped ncclBroadcast of the pruned row blocks" % world),
                       "l2": "no flush: every iteration reads the previous iteration's output; operands exceed L2 from scale 18 on"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}


def run_reference_rmcl(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import sparse_matrix_with_flops_b200 as smf  # generator library only (host code)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    A, desc, small = make_rmcl_workload(smf, args.workload)
    As, desc_s, _ = make_rmcl_workload(smf, small)
    M = ol.from_csr(As)
    kind = "reference" if os.path.exists(ol.REF_SO) else "port"
    iters = args.rmcl_iters

    def one():
        if kind == "reference":
            with quiet_stdout():
                return ol.r_rmcl_iter(M, M, iters).ms
        t0 = time.perf_counter()
        ol.o_rmcl_iter(M, M, iters)
        return (time.perf_counter() - t0) * 1e3
    for _ in range(max(1, min(args.warmup, 1))):
        one()
    ms = float(np.mean([one() for _ in range(max(1, min(args.steps, 3)))]))
    # rate-vs-rate: the sample's products / second against the full workload's products per iteration
    T, prods = M, 0
    for _ in range(iters):
        prods += int(ol.o_flops_prefix(M, T)[-1])
        T = ol.o_rmcl_onestep(M, T)
    Mf, Tf, full = ol.from_csr(A), None, []
    full_p0 = int(ol.o_flops_prefix(Mf, Mf)[-1])   # iteration 0 of the full workload (later iterates need the GPU)
    pps = prods / (ms * 1e-3)
    val = pps / full_p0
    cores = ol.ref().ref_num_threads() if kind == "reference" else 1
    sample = ("%s: %d iterations, %d products in %.1f ms = %.3g products/s (%.3f iter/s there); value = that rate / "
              "the full workload's first-iteration products %d" % (desc_s, iters, prods, ms, pps, iters / (ms * 1e-3), full_p0))
    print(json.dumps({
        "impl": "reference", "metric": RMCL_METRIC, "value": val, "unit": RMCL_UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "timing": "omp_get_wtime around the reference's mtRmclIter (SOMP)"},
        "cpu_baseline": {"value": val, "unit": RMCL_UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": RMCL_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ---- GPU arm -----------------------------------------------------------------------------------

def kernel_bytes(kind, rows, products, nnzA, nnzC):
    """ALGORITHMIC bytes of one launch (SURVEY.md §8d per-unit figures,
    int32 index + fp64 value): numeric = A rows (12/entry + 4/row) + gathered B rows (12/product
    + 8 of rowptr per A entry) + C rows (12/entry + 4/row); symbolic reads indices only and
    writes one count per row."""
    if kind == "num":
        return 12 * nnzA + 4 * rows + 12 * products + 8 * nnzA + 12 * nnzC + 4 * rows
    return 4 * nnzA + 4 * rows + 4 * products + 8 * nnzA + 4 * rows


def run_b200(args):
    import torch
    import torch.distributed as dist
    import sparse_matrix_with_flops_b200 as smf
    from sparse_matrix_with_flops_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local)
    # host block cache of the e2e leg: large enough to hold one step's result blocks of this rank
    # (the library's own default is min(RAM / 4, 64 GB) per process; read once, at first use)
    if "B200_HOST_CACHE_GB" not in os.environ:
        try:
            kb = int(next(l for l in open("/proc/meminfo") if l.startswith("MemTotal")).split()[1])
            os.environ["B200_HOST_CACHE_GB"] = "%.1f" % (min(0.5 * kb / 2**20, 160.0) / world)
        except Exception:  # noqa: BLE001
            pass
    if world > 1:
        # (NCCL prints its version banner on stdout when the first communicator is created)
        with quiet_stdout():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    smf.init(local)  # raises without a CUDA device: there is no CPU fallback
    lib = _lib.load()
    sp = C.c_void_p()
    _lib.check(lib.b200_stream(C.byref(sp)))
    stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env = {"smf": smf, "lib": lib, "stream": stream, "barrier": barrier, "rank": rank, "world": world,
           "local": local}
    if args.workload.startswith("rmcl-"):
        line = rmcl_measure(args, env, args.workload, args.steps, args.warmup, not args.no_cpu, not args.no_e2e)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            smf.comm_destroy()
            dist.destroy_process_group()
        return

    A, desc = make_workload(smf, args.workload)
    dA = A.toGpuCSR()
    prefix = smf.flops_prefix(dA, dA)
    P = int(prefix[-1])
    # Row blocks of equal COST: products plus a fixed charge for every row that goes through the
    # CTA-per-row bitmap kernels.  Equal products alone (the reference's thread partition,
    # arrayEqualPartition64 on the flops prefix) leaves the rank that holds the long tail of
    # lighter rows 1.7x slower than the rank that holds the hubs (measured at N=4: 50 / 63 / 75 /
    # 84 ms per rank); a charge of 32K products per heavy row gives 69 / 68 / 67 / 67 ms.
    # (b200_cost_prefix: the library's own cost model, the same the sharded rMCL loop cuts with)
    cost_prefix = smf.cost_prefix(dA, dA, args.row_charge)
    ends = smf.arrayEqualPartition64(cost_prefix, world)
    lo, hi = int(ends[rank]), int(ends[rank + 1])

    acc = {"sym": np.zeros(16), "num": np.zeros(16), "launches": 0, "last": None,
           "phases": np.zeros(5)}

    def step(record):
        dC, st = smf.gpuSpMMWrapper(dA, dA, lo, hi, want_stats=True)
        dC.deviceDispose()
        if record:
            acc["sym"] += np.array(st["ms_sym_bin"])
            acc["num"] += np.array(st["ms_num_bin"])
            acc["launches"] += st["launches"]
            acc["phases"] += np.array([st["ms_total"], st["ms_flops"], st["ms_symbolic"],
                                       st["ms_numeric"], st["ms_other"]])
            acc["last"] = st

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(True)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total, float(acc["launches"])], dtype=torch.float64, device="cuda")
    rank_ms = [ms_total / args.steps]
    if world > 1:
        allms = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(allms, t[:1].clone())
        rank_ms = [float(x[0]) / args.steps for x in allms]
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_total, launches = float(tmax[0]), int(t[1])
    else:
        launches = int(t[1])
    ms_step = ms_total / args.steps
    value = 2.0 * P / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (rank 0's launches) -----------------------------
    st = acc["last"]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    cands = []
    fused = st["bins_rows"][FUSED_BIN] > 0
    for b, nm in SYM_KERNELS.items():
        if acc["sym"][b] > 0:
            if b == 3 and fused:   # (+ the symbolic retry of the rows that did not fit 128 columns)
                nm = "k_num_warp_fused<128>"
            cands.append((acc["sym"][b] / args.steps, "sym", b, nm))
    for b, nm in NUM_KERNELS.items():
        if acc["num"][b] > 0:
            if b == 5 and st.get("part_kernel"):
                nm = "k_num_bitmap_part"
            cands.append((acc["num"][b] / args.steps, "num", b, nm))
    cands.sort(reverse=True)
    kms, kind, b, kname = cands[0]
    if kind == "num":
        kb = kernel_bytes("num", st["bins_rows"][b], st["num_bin_products"][b], st["num_bin_nnzA"][b],
                          st["num_bin_nnzC"][b])
    elif b in (6, 7):   # rows finished on chip in the symbolic phase: a numeric kernel in all but name
        kb = kernel_bytes("num", st["sym_bin_rows"][b], st["sym_bin_products"][b], st["sym_bin_nnzA"][b],
                          st["num_bin_nnzC"][b + 1])
    elif b == 3 and fused:   # likewise; the bytes of the rows it finished
        kb = kernel_bytes("num", st["bins_rows"][FUSED_BIN], st["num_bin_products"][FUSED_BIN],
                          st["num_bin_nnzA"][FUSED_BIN], st["num_bin_nnzC"][FUSED_BIN])
    else:
        kb = kernel_bytes("sym", st["sym_bin_rows"][b], st["sym_bin_products"][b], st["sym_bin_nnzA"][b], 0)
    achieved = kb / (kms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:   # the ncu capture is of the whole product on one GPU
        traffic = json.load(open(tpath)).get(args.workload, {}).get(kname)
    nnzC = st["nnz_out"] if world == 1 else None
    step_bytes = None
    if world == 1:
        step_bytes = kernel_bytes("num", A.rows, P, A.nnz, st["nnz_out"])
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": kms, "kernel_share_of_step": kms / ms_step,
                "kernel_algorithmic_bytes": kb,
                "step_algorithmic_bytes": step_bytes,
                "step_frac": (step_bytes / (ms_step * 1e-3) / 1e9 / peak) if step_bytes else None,
                "kernels_ms": {nm: round(ms_, 4) for ms_, _, _, nm in cands},
                "phases_ms": dict(zip(["call", "flops_binning", "symbolic", "numeric", "scans_alloc"],
                                      [round(float(x) / args.steps, 3) for x in acc["phases"]]))}

    # ---- e2e through the host-buffer C-ABI ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, smf, lib, A, rank, world, barrier, lo, hi, P)

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        arm = CpuArm(A)
        arm.step_ms()
        ms = min(arm.step_ms(), arm.step_ms())
        cpu = {"value": arm.gflops(ms), "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
               "sample": arm.sample, "ms": ms}

    # ---- the second half of the BASELINE metric: rMCL iterations / second, same process and GPUs
    rmcl = None
    if args.rmcl_leg != "none":
        dA.deviceDispose()
        dA = None
        try:
            # (two warm-up loops: the first grows the arena and the memory pool to their working sizes)
            r = rmcl_measure(args, env, args.rmcl_leg, max(1, min(args.steps, 2)), 2, not args.no_cpu,
                             not args.no_e2e)
            rmcl = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "config",
                                      "roofline", "cpu_baseline", "e2e", "gpu_launches")}
        except Exception as e:  # noqa: BLE001 — the headline line must still be printed
            rmcl = {"metric": RMCL_METRIC, "value": None, "error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "rows": A.rows, "nnzA": A.nnz, "products": P, "nnzC": nnzC,
                       "partition": "cost-balanced contiguous row blocks (products + %d per heavy row), 1 per GPU" % args.row_charge,
                       "rank_ms_per_step": [round(x, 3) for x in rank_ms],
                       "l2": "no flush: inputs (%.0f MB) and output exceed the %d MB L2" % (
                           (12 * A.nnz + 8 * A.rows) / 1e6, L2_BYTES // (1024 * 1024))},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "rmcl": rmcl,
        }
        print(json.dumps(line), flush=True)
    if dA is not None:
        dA.deviceDispose()
    if world > 1:
        if env.get("comm"):
            smf.comm_destroy()
        dist.destroy_process_group()


def run_e2e(args, smf, lib, A, rank, world, barrier, lo, hi, P):
    """The reference-facing call with HOST buffers, every copy inside the timed region:
    b200_spgemm_csr_stream = flops_omp_CSR_SpMM's operands (host int CSR of A, twice) in, the
    product out as malloc'd host int CSR row blocks handed to a callback that owns and frees
    them.  nnz(C) = 9.7e9 exceeds the reference's `int` CSR, so the library cuts row blocks where
    the intermediate-product prefix reaches 2e9 (an upper bound of a block's nnz)
    and overlaps the download of block b with the computation of block b+1.  With N > 1 ranks
    each rank walks its own row range with the upload / row-block product / row-block download
    sequence instead (same buffers, no overlap)."""
    from sparse_matrix_with_flops_b200 import _lib
    ip, dp = _lib.c_int_p, _lib.c_double_p
    h2d = d2h = nblk = 0
    same = lo == 0 and hi == A.rows
    # this rank's row block of A as a CSR of its own (row offsets from 0, its slice of JA / A)
    z0, z1 = int(A.rowPtr[lo]), int(A.rowPtr[hi])
    rp = np.ascontiguousarray(A.rowPtr[lo:hi + 1] - A.rowPtr[lo]).astype(np.int32)
    blkJ, blkV = A.colInd[z0:z1], A.values[z0:z1]

    @_lib.block_fn
    def sink(_user, r0, r1, IC, JC, Cv, nnzC):
        nonlocal d2h, nblk
        d2h += 4 * (r1 - r0 + 1) + 12 * nnzC
        nblk += 1
        for p in (IC, JC, Cv):                               # the caller owns the block
            lib.b200_host_free(C.cast(p, C.c_void_p))
        return 0

    def one_streamed():
        nonlocal h2d, d2h, nblk
        d2h = nblk = 0
        h2d = 4 * (hi - lo + 1) + 12 * (z1 - z0) + (0 if same else 4 * (A.rows + 1) + 12 * A.nnz)
        if same:   # the very same arrays on both sides: the library uploads them once
            _lib.check(lib.b200_spgemm_csr_stream(
                A.rowPtr.ctypes.data_as(ip), A.colInd.ctypes.data_as(ip), A.values.ctypes.data_as(dp), A.nnz,
                A.rowPtr.ctypes.data_as(ip), A.colInd.ctypes.data_as(ip), A.values.ctypes.data_as(dp), A.nnz,
                A.rows, A.cols, A.cols, 0, sink, None))
        else:
            _lib.check(lib.b200_spgemm_csr_stream(
                rp.ctypes.data_as(ip), blkJ.ctypes.data_as(ip), blkV.ctypes.data_as(dp), z1 - z0,
                A.rowPtr.ctypes.data_as(ip), A.colInd.ctypes.data_as(ip), A.values.ctypes.data_as(dp), A.nnz,
                hi - lo, A.rows, A.cols, 0, sink, None))

    # N > 1: CSR::toGpuCSR of the whole matrix, gpuSpMMWrapper on this rank's row blocks,
    # CSR::toCpuCSR of each block.  (The streamed call per rank was measured too, round 2, N = 2:
    # 8.5 GFLOP/s against 16.4 for this sequence — the ranks share the host's 16 cores, and the
    # streamed path's helper thread and staging copies of one rank then compete with the other
    # rank's kernels' host side)
    prefix = host_flops_prefix(A)
    mine = int(prefix[hi] - prefix[lo])
    nblk_seq = max(1, -(-mine // 2_000_000_000))
    cuts = smf.arrayEqualPartition64((prefix[lo:hi + 1] - prefix[lo]).astype(np.int64), nblk_seq) + lo

    def one_sequential():
        nonlocal h2d, d2h, nblk
        dA = A.toGpuCSR()
        h2d = 4 * (A.rows + 1) + 12 * A.nnz
        d2h = nblk = 0
        for b in range(nblk_seq):
            r0, r1 = int(cuts[b]), int(cuts[b + 1])
            if r1 <= r0:
                continue
            dC = smf.gpuSpMMWrapper(dA, dA, r0, r1)
            IC, JC, Cv, nnzC = ip(), ip(), dp(), C.c_int(0)
            _lib.check(lib.b200_csr_download_rows(dC.handle, 0, r1 - r0, C.byref(IC), C.byref(JC),
                                                  C.byref(Cv), C.byref(nnzC)))
            dC.deviceDispose()
            sink(None, r0, r1, IC, JC, Cv, nnzC.value)
        dA.deviceDispose()

    streamed = world == 1 if args.e2e_api == "auto" else args.e2e_api == "streamed"
    one = one_streamed if streamed else one_sequential
    api = ("b200_spgemm_csr_stream (host int CSR in, malloc'd host int CSR row blocks out through a callback)"
           if streamed else
           "b200_csr_upload + b200_spgemm_device_rows + b200_csr_download_rows per row block (host malloc'd int CSR in/out)")

    # Result blocks come back through b200_host_free; with the page-locked cache the blocks of the
    # warm-up step are registered once and every later download is one DMA into them.
    pin = not args.no_pin
    lib.b200_host_cache_pin(1 if pin else 0)
    # A rank whose call fails still takes part in every collective below (no deadlock); the leg
    # is then reported as failed instead of taking the whole bench line down.
    err = None
    try:
        one()  # warm-up (page-faults the pools, fills the host block cache, loads the kernels)
    except Exception as e:  # noqa: BLE001
        err = repr(e)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        if err is None:
            try:
                one()
            except Exception as e:  # noqa: BLE001
                err = repr(e)
    barrier()
    sec = (time.perf_counter() - t0) / args.e2e_steps
    hits, miss, direct, pinned = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_int()
    lib.b200_host_cache_stats(C.byref(hits), C.byref(miss), C.byref(direct), C.byref(pinned))
    cache = {"mode": "page-locked blocks (b200_host_cache_pin)" if pin else "malloc blocks (default)",
             "limit_gb": os.environ.get("B200_HOST_CACHE_GB"), "hits": hits.value, "misses": miss.value,
             "direct_downloads": direct.value, "page_locked_blocks_rank0": pinned.value}
    lib.b200_host_cache_drop()     # un-registers: the legs that follow get the host memory back
    lib.b200_host_cache_pin(0)
    import torch
    import torch.distributed as dist
    t = torch.tensor([sec, 1.0 if err else 0.0], dtype=torch.float64, device="cuda")
    io = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(io, op=dist.ReduceOp.SUM)
    if float(t[1]) > 0:
        return {"value": None, "unit": UNIT, "error": err or "another rank failed"}
    sec = float(t[0])
    return {"value": 2.0 * P / sec / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(io[0]),
            "d2h_bytes_per_step": int(io[1]), "ms_per_step": sec * 1e3, "steps": args.e2e_steps,
            "warmup": 1, "row_blocks": nblk, "host_block_cache": cache,
            "api": api + ", wall clock incl. H2D + D2H"}


def main():
    args = parse_args()
    if args.impl == "reference":
        (run_reference_rmcl if args.workload.startswith("rmcl-") else run_reference)(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
