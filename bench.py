#!/usr/bin/env python
"""bench.py -- SpMV throughput of the B200-native engine on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--no-extra]

One JSON line on stdout (rank 0).  The workload of EVERY GPU count is BASELINE configs[4], the one the metric
("... at 1/2/4/8 B200") is quoted on: CSR SpMV on the 27-point stencil of a 512^3 grid (134 217 728 rows,
3 609 741 304 non-zeros, 46.0 GB of algorithmic bytes), which fits one GPU.  A "step" is one product over the
whole matrix.

N = 1   y += A*x (the reference's Kernel::run) through spmvb200_spmv, inputs resident in HBM, K launches timed
        with CUDA events.  `value` = effective GB/s = (matrix_size + x_size + y_size) / t with the sizes the
        reference prints.  "targets" carries BASELINE's 70 %-of-8 TB/s cases (config 1, 2D 5-point 1000x1000, CSR
        and ELL) pipelined, isolated and L2-warm; "formats" the other formats / configurations (COO, hybrid,
        configs 2-4), each with its own roofline fraction.  Small workloads rotate over enough independent copies
        that neither the matrix nor x + y of a launch can be found in L2.
N > 1   the same matrix row-partitioned over N ranks (strong scaling): x_(k+1) = alpha*A*x_k, a step = the exchange
        of x between ranks + the SpMV, executed below the C ABI (spmvb200_dist_*, NCCL); torch.distributed only
        carries the NCCL id.  N = 2 adds BASELINE configs[3] (hybrid ELL+COO, R-MAT 2^26 x 32) as "c4_hyb".
        `--workload c4_hyb` makes that the headline instead.
Every arm checks its result before it times anything ("parity": sampled rows against the closed form of the
stencil / a numpy row product for R-MAT) and exits non-zero if the check fails.
--impl reference   the reference's own OpenMP kernels (oracle/_ref, compiled from the unmodified
        reference sources) on this box's host cores, same metric.  Config 5 exceeds the reference's int32
        sizes: its CPU sample is a 512x512x24 slab of the same operator (stated in the line).

Everything under oracle/ is used here only as the CPU baseline and never on the GPU path.
"""
from __future__ import annotations

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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "spmv_effective_bandwidth"
UNIT = "GB/s"
NOMINAL_HBM_GBS = 8000.0  # the "8 TB/s HBM3e roofline" BASELINE.json's metric is normalised to
C4_ROW_WEIGHT = 0.0       # per-row cost (in entries) of the configs[3] partition; see measure_c4_partitioned


def measured_peak():
    """HBM copy bandwidth measured by the driver on this pool (MEASURED_PEAKS.json), else the recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def ncu_traffic(key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            v = json.load(f).get(key)
            return int(v["read"] + v["write"]) if isinstance(v, dict) else v
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# clocks / throttle reasons during the measurement (NVML, the data nvidia-smi prints)
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.004):
        self.index, self.period = index, period_s
        self.samples = []  # (t, sm_mhz, reasons_mask, util)
        self.marks = {}
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover - depends on the box
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                util = int(nv.nvmlDeviceGetUtilizationRates(self.h).gpu)
                self.samples.append((time.perf_counter(), mhz, mask, util))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def mark(self, name):
        self.marks[name] = time.perf_counter()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1.0)

    def summary(self, t0=None, t1=None):
        if not self.ok:
            return self._nvidia_smi_once()
        sel = [s for s in self.samples if (t0 is None or s[0] >= t0) and (t1 is None or s[0] <= t1)]
        window = "timed region"
        if len(sel) < 3:
            sel, window = list(self.samples), "whole measurement (timed region shorter than 3 samples)"
        busy = [s for s in sel if s[3] > 0 or not (s[2] & 0x1)] or sel
        if not busy:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        mask = 0
        for s in busy:
            mask |= s[2]
        reasons = sorted(n for b, n in self.REASONS.items() if mask & b and n != "gpu_idle")
        return {"sm_mhz": float(np.median([s[1] for s in busy])), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(busy), "window": window}

    def _nvidia_smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(v) for v in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "window": "single nvidia-smi query"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------

# name -> (family, parameters, format, coo mode, description)
WORKLOADS = {
    "c1_csr": ("stencil", (0, 1000, 1000, 1), "CSR", 0, "CSR, 2D 5-point Poisson 1000x1000 (config 1)"),
    "c1_ell": ("stencil", (0, 1000, 1000, 1), "ELL", 0, "ELL, 2D 5-point Poisson 1000x1000 (config 1 matrix)"),
    "c1_coo": ("stencil", (0, 1000, 1000, 1), "COO", 0, "COO, 2D 5-point Poisson 1000x1000 (config 1 matrix)"),
    "c1_hyb": ("stencil", (0, 1000, 1000, 1), "HYB", 0, "hybrid, 2D 5-point Poisson 1000x1000 (config 1 matrix)"),
    "c2_ell": ("stencil", (1, 128, 128, 128), "ELL", 0, "ELL, 3D 7-point Poisson 128^3 (config 2)"),
    "c2_csr": ("stencil", (1, 128, 128, 128), "CSR", 0, "CSR, 3D 7-point Poisson 128^3 (config 2 matrix)"),
    "c3_coo": ("rmat", (24, 16, 0x5EED0003), "COO", 0, "COO (segmented), R-MAT 2^24 x 16 (config 3)"),
    "c3_coo_atomic": ("rmat", (24, 16, 0x5EED0003), "COO", 1, "COO (atomic), R-MAT 2^24 x 16 (config 3 matrix)"),
    "c4_hyb": ("rmat", (26, 32, 0x5EED0004), "HYB", 0, "hybrid, R-MAT 2^26 x 32 (config 4)"),
    "c5_csr": ("stencil", (2, 512, 512, 512), "CSR", 0, "CSR, 3D 27-point 512^3 (config 5, one GPU)"),
    # reduced sizes (profiling, quick checks)
    "c5s_csr": ("stencil", (2, 256, 256, 256), "CSR", 0, "CSR, 3D 27-point 256^3 (config 5 at 1/8 size)"),
    "c3s_coo": ("rmat", (21, 16, 0x5EED0003), "COO", 0, "COO (segmented), R-MAT 2^21 x 16 (config 3 at 1/8 size)"),
    "c4s_hyb": ("rmat", (22, 32, 0x5EED0004), "HYB", 0, "hybrid, R-MAT 2^22 x 32 (config 4 at 1/16 size)"),
}


def make_workload(sp, name: str):
    """Returns (factory() -> DeviceMatrix, description)."""
    family, prm, fmt, mode, desc = WORKLOADS[name]
    f = getattr(sp, fmt)
    if family == "stencil":
        kind, nx, ny, nz = prm
        return (lambda: sp.generators.stencil(kind, nx, ny, nz, f)), desc
    scale, ef, seed = prm
    return (lambda: sp.generators.rmat(scale, ef, seed, fmt=f, coo_mode=mode)), desc


# ------------------------------------------------------------------------------------------------
# parity before timing: sampled rows of one product against arithmetic done here in numpy
# ------------------------------------------------------------------------------------------------

def x_pattern(j):
    """x_j = 1 + (j mod 7)/8: exactly representable, so stencil rows have exact closed-form values."""
    return 1.0 + (np.asarray(j, dtype=np.int64) % 7) / 8.0


STENCIL_OFFSETS = {
    0: ([(0, -1, 0), (0, 0, -1), (0, 0, 1), (0, 1, 0)], 4.0),
    1: ([(-1, 0, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (0, 1, 0), (1, 0, 0)], 6.0),
    2: ([(dz, dy, dx) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dz, dy, dx) != (0, 0, 0)], 26.0),
}


def stencil_rows_closed_form(kind, nx, ny, nz, rows):
    """(A x)_i for x = x_pattern and the rows given (global indices): diag*x_i - sum of the neighbours inside the grid.
    Every term is a multiple of 1/8 below 2^10, so the sums are exact in fp64 in any order."""
    r = np.asarray(rows, dtype=np.int64)
    ix, iy, iz = r % nx, (r // nx) % ny, r // (nx * ny)
    offs, diag = STENCIL_OFFSETS[kind]
    y = diag * x_pattern(r)
    for dz, dy, dx in offs:
        ok = (ix + dx >= 0) & (ix + dx < nx) & (iy + dy >= 0) & (iy + dy < ny) & (iz + dz >= 0) & (iz + dz < nz)
        y -= np.where(ok, x_pattern(r + (dz * ny + dy) * nx + dx), 0.0)
    return y


def sample_rows(row_begin, row_end, plane, count, seed):
    """Rows to check: everything within `plane` rows of either end of [row_begin, row_end) (the rows whose columns
    cross a partition boundary) plus `count` random rows."""
    n = row_end - row_begin
    if n <= 2 * plane + count:
        return np.arange(row_begin, row_end, dtype=np.int64)
    rng = np.random.default_rng(seed)
    mid = rng.integers(row_begin + plane, row_end - plane, size=count, dtype=np.int64)
    return np.unique(np.concatenate([np.arange(row_begin, row_begin + plane), mid, np.arange(row_end - plane, row_end)]))


def parity_stencil(kind, nx, ny, nz, row_begin, y_local, seed=7):
    """Exact comparison of a rank's rows of A*x_pattern with the closed form on sampled rows."""
    plane = nx * ny if nz > 1 else nx
    rows = sample_rows(row_begin, row_begin + len(y_local), min(plane + nx + 2, len(y_local)), 1 << 20, seed)
    ref = stencil_rows_closed_form(kind, nx, ny, nz, rows)
    got = y_local[rows - row_begin]
    bad = int(np.count_nonzero(got != ref))
    scale = np.maximum(np.abs(ref), 1.0)
    return {"rows_checked": int(len(rows)), "max_rel": float(np.max(np.abs(got - ref) / scale)) if len(rows) else 0.0,
            "bad_rows": bad, "ok": bad == 0, "against": "closed form of the stencil, x_j = 1 + (j mod 7)/8, exact"}


def parity_csr_rows(rp, col, val, y_rows):
    """|y - sum a_ij x_j| <= 1e-12 * sum |a_ij x_j| per row (BASELINE tolerance) for CSR rows held on the host."""
    p = val * x_pattern(col)
    rp = np.asarray(rp, dtype=np.int64)
    nonempty = rp[1:] > rp[:-1]
    ref = np.zeros(len(rp) - 1)
    bound = np.zeros(len(rp) - 1)
    if len(p):
        # reduceat over the starts of the NON-EMPTY rows only (strictly increasing, the last one sums to the end): with
        # empty rows in the index list reduceat would cut the preceding row short
        idx = rp[:-1][nonempty]
        ref[nonempty] = np.add.reduceat(p, idx)
        bound[nonempty] = np.add.reduceat(np.abs(p), idx)
    err = np.abs(y_rows - ref)
    bad = int(np.count_nonzero(err > 1e-12 * bound))
    rel = float(np.max(err / np.maximum(bound, 1e-300))) if len(err) else 0.0
    return bad, rel


def parity_single(sp, name, A):
    """One product on the headline matrix with x_pattern, checked before anything is timed."""
    family, prm = WORKLOADS[name][0], WORKLOADS[name][1]
    n = A.rows
    y = A * x_pattern(np.arange(A.columns))
    if family == "stencil":
        return parity_stencil(*prm, 0, y)
    # R-MAT: the rows of two row ranges (the hub rows at the top, a range in the middle) regenerated as CSR
    scale, ef, seed = prm
    bad, rel, checked = 0, 0.0, 0
    for r0 in (0, n // 2 + 12345):
        r1 = min(n, r0 + 4096)
        blk = sp.generators.rmat(scale, ef, seed, fmt=sp.CSR, row_begin=r0, row_end=r1).export()
        b, r = parity_csr_rows(blk["row_ptr"], blk["column_index"], blk["value"], y[r0:r1])
        bad, rel, checked = bad + b, max(rel, r), checked + (r1 - r0)
    return {"rows_checked": checked, "max_rel": rel, "bad_rows": bad, "ok": bad == 0,
            "against": "numpy row products of regenerated CSR rows, |err| <= 1e-12 * sum|a_ij x_j| per row"}


# ------------------------------------------------------------------------------------------------
# device-resident timing of one workload
# ------------------------------------------------------------------------------------------------

def l2_cold_copies(A, l2_bytes, free_bytes):
    """Copies to rotate over so that no launch finds its operands in L2: the cycle holds >= 3 x L2 of matrix data AND
    >= 2 x L2 of x + y (the matrix streams are evict-first, so x and y are what could survive a short cycle)."""
    inf = A.info
    B = A.algorithmic_bytes()
    if B >= 3 * l2_bytes:
        return 1
    xy = int(inf.x_size + inf.y_size)
    want = max(int(math.ceil(3.0 * l2_bytes / B)), int(math.ceil(2.0 * l2_bytes / max(xy, 1))))
    fit = max(1, int(0.5 * free_bytes / max(int(inf.device_bytes), 1)))
    return max(1, min(want, fit, 40))


def free_device_bytes():
    import ctypes
    try:
        cudart = ctypes.CDLL("libcudart.so")
        free, total = ctypes.c_size_t(), ctypes.c_size_t()
        if cudart.cudaMemGetInfo(ctypes.byref(free), ctypes.byref(total)) == 0:
            return int(free.value)
    except OSError:
        pass
    return 160 << 30


def measure_device(sp, name, steps, warmup, l2_bytes, peak, isolated=True, warm=False):
    """Device-resident timing of one workload.  pipelined: K back-to-back launches between one event pair;
    isolated: one event pair per launch (the barrier-run-barrier protocol of profile_kernel_run with events for the
    clock); warm: the same matrix every launch (the reference protocol without --flush-caches: operands stay cached)."""
    make, desc = make_workload(sp, name)
    A = make()
    B = A.algorithmic_bytes()
    inf = A.info
    copies = l2_cold_copies(A, l2_bytes, free_device_bytes())
    mats = [A] + [make() for _ in range(copies - 1)]
    total_ms, _ = sp.time_rotating(mats, steps, warmup, False)
    t = total_ms * 1e-3 / steps

    def rates(ms):
        tt = ms * 1e-3
        return {"ms": ms, "gbs": B / tt / 1e9, "frac_of_8TBs": B / tt / 1e9 / NOMINAL_HBM_GBS, "frac_of_measured_peak": B / tt / 1e9 / peak}

    res = {
        "workload": name, "description": desc, "kernel": A.kernel_name,
        "rows": int(inf.rows), "nonzeros": int(inf.num_entries), "algorithmic_bytes": int(B),
        "ms_per_step": total_ms / steps, "gbs": B / t / 1e9, "gflops": 2.0 * inf.num_entries / t / 1e9,
        "frac_of_8TBs": B / t / 1e9 / NOMINAL_HBM_GBS, "frac_of_measured_peak": B / t / 1e9 / peak,
        "l2_cold_copies": copies, "xy_bytes_in_cycle": int(copies * (inf.x_size + inf.y_size)),
        "traffic": ncu_traffic(name), "resident_bytes": int(inf.device_bytes),
    }
    if inf.format in (sp.HYB, sp.ELL):
        res["ell_row_length"] = int(inf.ell_row_length)
    if inf.format == sp.HYB:
        res["num_coo_entries"] = int(inf.num_coo_entries)
    if isolated:
        _, per = sp.time_rotating(mats, min(steps, 500), 3, True)
        res["isolated"] = rates(float(np.mean(per)))
        res["isolated"]["ms_median"] = float(np.median(per))
    if warm:
        ms = A.time(reps=min(steps, 200), warmup=10)
        res["l2_warm_isolated"] = rates(float(np.median(ms)))
    return res, mats


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own kernels on this box's host cores
# ------------------------------------------------------------------------------------------------

def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


CPU_SAMPLES = {
    "c2_ell": (1, (128, 128, 128), "ell"), "c1_csr": (0, (1000, 1000, 1), "csr"), "c1_ell": (0, (1000, 1000, 1), "ell"),
    "c1_coo": (0, (1000, 1000, 1), "coo"), "c1_hyb": (0, (1000, 1000, 1), "hybrid"), "c2_csr": (1, (128, 128, 128), "csr"),
    # config 5 exceeds the reference's int32 sizes: the baseline sample is a 512x512x24 slab of the same operator
    "c5_csr": (2, (512, 512, 24), "csr"),
}


def cpu_baseline(workload: str, reps: int, warmup: int = 1, threads: int = 0):
    """Times oracle/_ref (the unmodified reference library) -- or the oracle port if it is missing --
    with the protocol of profile_kernel_run, on the SAME operator as the GPU workload."""
    from oracle.generators_ref import stencil_entries
    from oracle.oracle import Oracle, Ref
    T = threads or host_threads()
    kind, dims, fmt = CPU_SAMPLES[workload]
    i, j, a = stencil_entries(kind, *dims)
    n = dims[0] * dims[1] * dims[2]
    sample = f"{fmt.upper()} y+=A*x on the {dims[0]}x{dims[1]}x{dims[2]} stencil matrix ({n} rows, {len(i)} nnz), " \
             f"x=1, 1 warm-up + {reps} timed runs, barrier/steady_clock/barrier (profile-kernel.cpp:137-179)"
    if workload == "c5_csr":
        sample += "; a SLAB of config 5's 512^3 operator: the full matrix (3.6 G non-zeros) does not fit the reference's int32 sizes"
    if Ref.available():
        m = Ref().from_entries(n, n, i, j, a)
        A = m.convert(fmt)
        size = A.size if fmt != "hybrid" else 12 * A.num_ell_entries + 16 * A.num_coo_entries
        ns = m.time(threads=T, reps=reps, pin=True)
        kind_s = "reference"
    else:  # the oracle port, really threaded
        orc = Oracle()
        x = np.ones(n)
        if fmt == "ell":
            A = orc.ell(n, n, i, j, a); size = A.size; run = lambda: orc.ell_spmv(A, x, threads=T)
        else:
            A = orc.csr(n, n, i, j, a); size = A.size; run = lambda: orc.csr_spmv(A, x, threads=T)
        run()
        ns = []
        for _ in range(reps):
            t0 = time.perf_counter(); run(); ns.append((time.perf_counter() - t0) * 1e9)
        ns = np.array(ns)
        kind_s = "port"
    B = size + 16 * n
    t_med = float(np.median(ns)) * 1e-9
    return {"value": B / t_med / 1e9, "unit": UNIT, "cores": T, "kind": kind_s, "sample": sample,
            "ms_median": t_med * 1e3, "ms_min": float(np.min(ns)) * 1e-6, "algorithmic_bytes": int(B),
            "gflops": 2.0 * len(i) / t_med / 1e9}


# ------------------------------------------------------------------------------------------------
# arms
# ------------------------------------------------------------------------------------------------

def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload or "c5_csr"
    if wl not in CPU_SAMPLES:
        # R-MAT configs exceed what the reference's int32 conversion can hold in host memory here
        wl = "c5_csr"
    t0 = time.perf_counter()
    cb = cpu_baseline(wl, reps=max(args.steps, 1), warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_median"], "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "host_threads": cb["cores"], "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops": cb["gflops"], "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_single_gpu(args):
    import spmv_cache_trace_b200 as sp

    if sp.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback")
    props = sp.device_props(0)
    peak, peak_src = measured_peak()
    wl = args.workload or "c5_csr"
    sampler = ClockSampler(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0] or 0))
    sampler.start()

    # ---- headline: device-resident -----------------------------------------------------------------------------------
    make, desc = make_workload(sp, wl)
    A = make()
    B = A.algorithmic_bytes()
    inf = A.info
    parity = parity_single(sp, wl, A)
    if not parity["ok"]:
        print(json.dumps({"metric": METRIC, "error": "parity check failed", "parity": parity}), flush=True)
        return 3
    A.fill_x(1.0)  # the reference's operands: x = 1, y accumulating from 0 (csr-spmv.cpp:35-36)
    A.fill_y(0.0)
    copies = l2_cold_copies(A, props["l2_bytes"], free_device_bytes())
    mats = [A] + [make() for _ in range(copies - 1)]
    # W untimed warm-up steps as asked.  A launch of the small workloads lasts 13-36 us, so W of them do not bring the
    # clocks up after the idle time of matrix generation: those get extra untimed launches BEFORE the W (reported).
    extra_warm = 0 if copies == 1 else 300
    if extra_warm:
        sp.time_rotating(mats, extra_warm, 0, False)
    sp.time_rotating(mats, max(args.warmup, 1), 0, False)
    launches0 = sp.launch_count()
    sampler.mark("t0"); t0 = time.perf_counter()
    total_ms, _ = sp.time_rotating(mats, args.steps, 0, False)  # the timed region: EXACTLY K steps
    t1 = time.perf_counter(); sampler.mark("t1")
    gpu_launches = sp.launch_count() - launches0
    _, per = sp.time_rotating(mats, args.steps, 0, True)  # same K steps again, one event pair per launch
    per = per if per is not None else np.array([total_ms / args.steps])
    t_step = total_ms * 1e-3 / args.steps
    value = B / t_step / 1e9
    k_ms = total_ms / args.steps
    iso_ms = float(np.mean(per))
    achieved = B / (k_ms * 1e-3) / 1e9
    kernel_name = A.kernel_name
    resident = int(A.info.device_bytes)
    index_runs = None
    if inf.format == sp.CSR and kernel_name == "csr_sliced_kernel":
        index_runs = {"active": bool(A.get_option("csr.index_runs_active")),
                      "column_indices_stored": int(A.get_option("csr.index_columns_stored")), "stored_entries": int(inf.num_entries),
                      "what": "slot-major copy of the sliced kernel: a 32-row slice whose entries lie on few diagonals stores its "
                              "slots by offset (column - row), one descriptor per slot instead of up to 32 column indices, and needs "
                              "no row_ptr; values untouched, same summation order, bit-identical results.  algorithmic_bytes keep "
                              "counting 4 B per column index and row_ptr like the reference's csr_matrix::size()"}

    # ---- end to end: host buffers through the C ABI ------------------------------------------------------------------
    e2e_steps = max(3, min(args.steps, 10 if copies == 1 else 50))
    xs = [sp.PinnedBuffer(int(inf.columns)) for _ in mats]
    ys = [sp.PinnedBuffer(int(inf.rows)) for _ in mats]
    for xb, yb in zip(xs, ys):
        xb.array[:] = 1.0
        yb.array[:] = 0.0
    e2e_ms = sp.time_host_rotating(mats, [b.array for b in xs], [b.array for b in ys], e2e_steps, 1)
    e2e_t = e2e_ms * 1e-3 / e2e_steps
    ycheck = float(np.abs(ys[0].array).max())  # the e2e result is read, not just timed
    sampler.stop()
    clocks = sampler.summary(t0, t1)
    if index_runs and index_runs["active"] and copies == 1:
        # the same K steps with every column index stored (the copy is rebuilt): what the index runs are worth
        A.set_option("csr.index_runs", -1)
        sp.time_rotating(mats, 2, 0, False)
        plain_ms, _ = sp.time_rotating(mats, args.steps, 0, False)
        index_runs["ms_per_step_with_every_index_stored"] = plain_ms / args.steps
        index_runs["traffic_with_every_index_stored"] = ncu_traffic(wl + "_plain_index")
        A.set_option("csr.index_runs", 0)
    traffic = ncu_traffic(wl)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "description": desc, "rows": int(inf.rows), "nonzeros": int(inf.num_entries),
                   "algorithmic_bytes": int(B), "semantics": "y += A*x (reference Kernel::run)",
                   "l2": (f"inputs ({B / 1e9:.1f} GB per launch) far larger than the {props['l2_bytes'] / 1e6:.0f} MB L2" if copies == 1 else
                          f"cold: launches rotate over {copies} independent copies of matrix, x and y "
                          f"({copies * B / 1e6:.0f} MB of matrix data, {copies * (inf.x_size + inf.y_size) / 1e6:.0f} MB of x+y per cycle "
                          f"vs {props['l2_bytes'] / 1e6:.0f} MB L2)"),
                   "extra_untimed_warmup_launches": extra_warm, "resident_bytes": resident,
                   "device": props["name"], "sm_count": props["sm_count"]},
        "gflops": 2.0 * inf.num_entries / t_step / 1e9,
        "frac_of_8TBs_nominal": value / NOMINAL_HBM_GBS,
        "parity": parity,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "dram_gbs": (traffic / (k_ms * 1e-3) / 1e9 if traffic else None),
                     "dram_frac_of_peak": (traffic / (k_ms * 1e-3) / 1e9 / peak if traffic else None),
                     "index_runs": index_runs, "kernel": kernel_name, "kernel_ms_mean": k_ms,
                     "isolated_launch_ms_mean": iso_ms, "isolated_launch_ms_median": float(np.median(per)),
                     "isolated_launch_gbs": B / (iso_ms * 1e-3) / 1e9, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(B),
                     "how": "CUDA events on the launching stream around the K launches of the timed region (kernel_ms_mean = "
                            "their average duration); isolated_launch_* = one event pair per launch in a second pass",
                     "note": "peak is the driver's COPY bandwidth (read+write); a read-dominated stream can exceed it.  achieved "
                             "= ALGORITHMIC bytes / time; when index_runs.active the kernel moves fewer bytes than that "
                             "(traffic, measured by ncu; dram_gbs = traffic / time is the physical DRAM rate)"},
        "e2e": {"value": B / e2e_t / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(8 * (inf.columns + inf.rows)),
                "d2h_bytes_per_step": int(8 * inf.rows), "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                "call": "spmvb200_spmv_host (pinned host x, y -> device, kernel, y -> host)", "max_abs_y": ycheck},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
    }
    del xs, ys, mats, A

    # ---- CPU baseline (rank 0, bounded sample) -----------------------------------------------------------------------
    if not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline(wl if wl in CPU_SAMPLES else "c5_csr", reps=10)
        except Exception as e:  # keep the GPU line even if the checker is unavailable
            line["cpu_baseline"] = {"error": repr(e)}

    # ---- BASELINE's 70 % targets and the other formats / configurations ----------------------------------------------
    if not args.no_extra:
        targets = {}
        for name in ("c1_csr", "c1_ell"):
            try:
                r, keep = measure_device(sp, name, 2000, 200, props["l2_bytes"], peak, isolated=True, warm=True)
                del keep
                targets[name] = {
                    "kernel": r["kernel"], "us_pipelined": r["ms_per_step"] * 1e3, "gbs_pipelined": r["gbs"],
                    "frac_of_8TBs_pipelined": r["frac_of_8TBs"], "us_isolated": r["isolated"]["ms"] * 1e3,
                    "frac_of_8TBs_isolated": r["isolated"]["frac_of_8TBs"],
                    "us_l2_warm_isolated": r["l2_warm_isolated"]["ms"] * 1e3,
                    "frac_of_8TBs_l2_warm_isolated": r["l2_warm_isolated"]["frac_of_8TBs"],
                    "l2_cold_copies": r["l2_cold_copies"], "xy_bytes_in_cycle": r["xy_bytes_in_cycle"],
                    "target": "BASELINE: >= 0.70 of 8 TB/s"}
            except Exception as e:
                targets[name] = {"error": str(e)}
        # the yardstick for the isolated figures: the same event-pair protocol around a plain device-to-device copy that
        # moves the same 80 MB of DRAM traffic (40 MB read + 40 MB written), L2-cold -- no kernel of ours involved
        try:
            cp = sp.time_copy(40_000_000, copies=8, reps=300, warmup=30)
            targets["isolated_copy_yardstick"] = {
                "what": "cudaMemcpyAsync device-to-device of 40 MB (80 MB of DRAM traffic, like one config-1 SpMV), event pair per copy, "
                        "rotating over 8 buffer pairs", "us_median": float(np.median(cp)) * 1e3, "us_min": float(np.min(cp)) * 1e3,
                "frac_of_8TBs": 80e6 / (float(np.median(cp)) * 1e-3) / 1e9 / NOMINAL_HBM_GBS}
        except Exception as e:
            targets["isolated_copy_yardstick"] = {"error": str(e)}
        line["targets"] = targets
        extra = []
        for name in ("c1_coo", "c1_hyb", "c2_ell", "c2_csr", "c3_coo", "c3_coo_atomic", "c4_hyb"):
            if name == wl:
                continue
            try:
                big = name in ("c3_coo", "c3_coo_atomic", "c4_hyb")
                r, keep = measure_device(sp, name, 20 if big else 2000, 3 if big else 200, props["l2_bytes"], peak, isolated=not big)
                del keep
                extra.append(r)
            except Exception as e:
                extra.append({"workload": name, "error": str(e)})
        line["formats"] = extra
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# N > 1: the row-partitioned mode (spmvb200_dist_* below the C ABI; torch.distributed carries the NCCL id)
# ------------------------------------------------------------------------------------------------

def _init_ranks():
    import torch
    import torch.distributed as dist

    import spmv_cache_trace_b200 as sp
    from spmv_cache_trace_b200 import distributed as D
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local_rank)
    sp.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = D.Comm.from_torch(dist, torch)
    else:
        comm = D.Comm.local(1, [local_rank])[0]
    return sp, D, dist, rank, world, local_rank, comm


def _e2e_host(sp, eng, rows, steps, alpha):
    """`steps` independent products through pinned host slices (spmvb200_dist_run_host); device ms per step."""
    xs = [sp.PinnedBuffer(rows) for _ in range(2)]
    ys = [sp.PinnedBuffer(rows) for _ in range(2)]
    rng = np.random.default_rng(3)
    for b in xs:
        b.array[:] = rng.random(rows) - 0.5
    eng.run_host([xs[i % 2].array for i in range(2)], [ys[i % 2].array for i in range(2)], alpha)  # warm-up
    ms = eng.run_host([xs[i % 2].array for i in range(steps)], [ys[i % 2].array for i in range(steps)], alpha)
    return ms / steps, float(np.abs(ys[0].array).max())


def measure_c5_partitioned(args, sp, D, comm, rank, world, sampler):
    n = int(os.environ.get("SPMV_BENCH_GRID", "512"))
    N = n ** 3
    starts = D.partition_rows_ref(N, world)
    s, e = int(starts[rank]), int(starts[rank + 1])
    nnz = (3 * n - 2) ** 3
    B = 4 * (N + 1) + 12 * nnz + 16 * N  # matrix_size + x_size + y_size, reference-equivalent
    # x_(k+1) = A x_k / 52: every eigenvalue of the 27-point operator (26 on the diagonal, -1 off it) lies within 52 of
    # zero, so the iterates stay finite however many steps are timed.
    SCALE = 1.0 / 52.0
    plans = [args.exchange] if getattr(args, "exchange", None) else ["allgather", "halo"]
    # every plan with both transports: NCCL kernels, and copy-engine pulls out of the owners' IPC-mapped buffers ("_peer")
    transports = [False, True] if world > 1 and os.environ.get("SPMV_PEER_COPY", "1") != "0" else [False]
    results, marks = {}, {}
    x3_first = None
    variants = [(m, t, False) for m in plans for t in transports]
    if len(transports) > 1 and "halo" in plans:
        variants.append(("halo", True, True))  # "_push": the kernels store the halo rows straight into the neighbours' buffers
    for mode, peer, push in variants:
        label = mode + ("_push" if push else "_peer" if peer else "")
        # every rank generates only its own rows of the 27-point operator, on its own GPU
        local = sp.generators.stencil(sp.STENCIL_3D27, n, n, n, fmt=sp.CSR, row_begin=s, row_end=e)
        eng = D.DistributedSpMV(comm, local, starts, mode=mode, consume_local=True, peer_copy=peer, peer_push=push)
        # parity before timing: one step from x_j = 1 + (j mod 7)/8, alpha = 1, every rank checks its rows exactly
        eng.set_x(x_pattern(np.arange(s, e)))
        eng.step(1.0)
        par = parity_stencil(2, n, n, n, s, eng.get_x(), seed=11 + rank)
        # ... and three steps back to back on the same exact data (multiples of 1/8 below 2^19: exact in any order), so that
        # steps 2 and 3 consume the halo the previous step delivered (copied, or stored by the fused push): every variant
        # must reproduce the first one's rows bit for bit
        eng.set_x(x_pattern(np.arange(s, e)))
        for _ in range(3):
            eng.step(1.0)
        x3 = eng.get_x()
        if x3_first is None:
            x3_first = x3
            par["three_steps"] = "reference variant"
        else:
            diff = int(np.count_nonzero(x3 != x3_first))
            par["bad_rows"] += diff
            par["three_steps"] = "identical to the first variant" if diff == 0 else f"{diff} rows differ from the first variant"
        par["bad_rows"] = int(comm.allreduce(par["bad_rows"], "sum"))
        par["rows_checked"] = int(comm.allreduce(par["rows_checked"], "sum"))
        par["max_rel"] = comm.allreduce(par["max_rel"], "max")
        par["ok"] = par["bad_rows"] == 0
        par["ranks"] = world
        if not par["ok"]:
            results[label] = {"parity": par}
            eng.destroy()
            continue
        eng.set_x(np.random.default_rng(1234 + rank).random(e - s) - 0.5)
        launches0 = sp.launch_count()
        marks[label] = [time.perf_counter(), None]
        ms = eng.time(args.steps, max(args.warmup, 3), SCALE)
        marks[label][1] = time.perf_counter()
        ms = comm.allreduce(ms, "max")  # device time, max over ranks
        t = ms * 1e-3 / args.steps
        inf = eng.info
        xnorm = math.sqrt(comm.allreduce(float(np.sum(eng.get_x() ** 2)), "sum"))
        results[label] = {"ms_per_step": t * 1e3, "gbs": B / t / 1e9, "gflops": 2.0 * nnz / t / 1e9,
                         "recv_bytes_per_step_per_rank": inf["recv_bytes_per_step"], "plan": inf["exchange"],
                         "transport": ("halo rows stored into the neighbours' IPC-mapped buffers by the SpMV kernel itself (fused push)"
                                       if push and inf["halo_push"] else "copy-engine pulls from IPC-mapped peer buffers" if peer else "NCCL"),
                         "row_blocks": eng.blocks(), "interior_rows": inf["interior_rows"],
                         "gpu_launches": int(sp.launch_count() - launches0), "launches_per_step": inf["launches_per_step"],
                         "resident_bytes_rank0": inf["device_bytes"], "x_norm": xnorm, "parity": par,
                         "kernel": max(eng.blocks(), key=lambda b: b[1] - b[0])[3]}
        e2e_steps = max(4, min(args.steps, 10))
        e2e_ms, ymax = _e2e_host(sp, eng, e - s, e2e_steps, SCALE)
        results[label]["e2e_ms_per_step"] = comm.allreduce(e2e_ms, "max")
        results[label]["e2e_steps"] = e2e_steps
        results[label]["e2e_max_abs_y"] = ymax
        eng.destroy()
        del eng, local
    return {"results": results, "marks": marks, "B": B, "N": N, "nnz": nnz, "grid": n, "scale": SCALE}


def measure_c4_partitioned(args, sp, D, comm, rank, world):
    """BASELINE configs[3]: rows cut into `world` blocks of equal non-zeros; every rank converts ITS rows to the hybrid
    format (the ELL part and the COO tail both follow the row owner); x is all-gathered between iterations.
    A power-law block references all of x, so the block is cut by COLUMNS: the entries that reference the rank's own
    slice of x run during the all-gather, the rest afterwards."""
    scale_log2 = int(os.environ.get("SPMV_BENCH_RMAT_SCALE", "26"))
    ef = int(os.environ.get("SPMV_BENCH_RMAT_EF", "32"))
    seed = 0x5EED0004
    N = 1 << scale_log2
    full = sp.generators.rmat(scale_log2, ef, seed, fmt=sp.CSR)  # every rank: the partition needs the global row_ptr
    # balanced COST: a row costs its entries + ROW_WEIGHT entries of per-row work (y traffic of two pieces, ELL padding, one
    # reduction per run); 0 = balanced non-zeros.  R-MAT's long rows come first, so with equal non-zeros the last rank holds
    # six times the rows of the first and is the slower one (`rank_compute_ms_without_exchange`).
    row_weight = float(os.environ.get("SPMV_C4_ROW_WEIGHT", str(C4_ROW_WEIGHT)))
    starts = sp.partition.rows_weighted(full, world, row_weight)
    s, e = int(starts[rank]), int(starts[rank + 1])
    block = full.row_block(s, e)
    nnz = full.num_entries
    del full
    # rows kept on the host for the parity check: the first rows of the block (rank 0: the hubs) and a range inside it
    samples = []
    for r0 in (0, max(0, (e - s) // 2 - 2048)):
        r1 = min(e - s, r0 + 4096)
        if r1 > r0:
            samples.append((r0, r1, block.row_block(r0, r1).export()))
    # the reference's hybrid of this rank's rows defines the bytes counted (one ELL part + one COO tail per rank)
    hinf = block.convert(sp.HYB).info
    mine = [12 * hinf.num_ell_entries + 16 * hinf.num_coo_entries, hinf.num_coo_entries, hinf.ell_row_length]
    per_rank = []
    for q in range(world):
        per_rank.append([int(comm.allreduce(float(v) if q == rank else 0.0, "sum")) for v in mine])
    B = sum(p[0] for p in per_rank) + 16 * N
    ALPHA = 1.0 / 8192.0  # keeps x_(k+1) = alpha A x_k finite: hub rows of the R-MAT matrix sum ~10^6 entries
    split = os.environ.get("SPMV_COLUMN_SPLIT", "1") != "0" and world > 1
    peer = world > 1 and os.environ.get("SPMV_PEER_COPY", "1") != "0" and os.environ.get("SPMV_C4_PEER", "1") != "0"
    eng = D.DistributedSpMV(comm, block, starts, mode="allgather", fmt=sp.HYB, column_split=split, overlap=False,
                            consume_local=True, peer_copy=peer)
    eng.set_x(x_pattern(np.arange(s, e)))
    eng.step(1.0)
    y = eng.get_x()
    bad, rel, checked = 0, 0.0, 0
    for r0, r1, blk in samples:
        b, r = parity_csr_rows(blk["row_ptr"], blk["column_index"], blk["value"], y[r0:r1])
        bad, rel, checked = bad + b, max(rel, r), checked + (r1 - r0)
    par = {"rows_checked": int(comm.allreduce(checked, "sum")), "bad_rows": int(comm.allreduce(bad, "sum")),
           "max_rel": comm.allreduce(rel, "max"), "ranks": world,
           "against": "numpy row products of sampled CSR rows, |err| <= 1e-12 * sum|a_ij x_j| per row"}
    par["ok"] = par["bad_rows"] == 0
    out = {"parity": par, "B": B, "N": N, "nnz": int(nnz), "starts": [int(v) for v in starts], "per_rank": per_rank,
           "split": split, "scale_log2": scale_log2, "ef": ef, "alpha": ALPHA, "row_weight": row_weight,
           "transport": "copy-engine pulls from IPC-mapped peer buffers" if peer else "NCCL"}
    if not par["ok"]:
        return out
    eng.set_x(np.random.default_rng(99 + rank).random(e - s) - 0.5)
    launches0 = sp.launch_count()
    t0 = time.perf_counter()
    ms = comm.allreduce(eng.time(args.steps, max(args.warmup, 3), ALPHA), "max")
    out["marks"] = (t0, time.perf_counter())
    inf = eng.info
    out.update(ms_per_step=ms / args.steps, gpu_launches=int(sp.launch_count() - launches0), recv_bytes=inf["recv_bytes_per_step"],
               blocks=eng.blocks(), x_norm=math.sqrt(comm.allreduce(float(np.sum(eng.get_x() ** 2)), "sum")),
               pieces=[{"rows": int(m.info.rows), "ell_row_length": int(m.info.ell_row_length), "num_coo_entries": int(m.info.num_coo_entries)}
                       for m in (eng.block_matrix(b) for b in range(inf["n_blocks"]))])
    e2e_steps = max(4, min(args.steps, 10))
    e2e_ms, _ = _e2e_host(sp, eng, e - s, e2e_steps, ALPHA)
    out["e2e_ms_per_step"] = comm.allreduce(e2e_ms, "max")
    # diagnostic: every rank's pieces timed alone, without the exchange (how well the equal-non-zeros cut balances the TIME)
    mine = 0.0
    for b in range(inf["n_blocks"]):
        total_ms, _ = sp.time_rotating([eng.block_matrix(b)], 5, 2, False)
        mine += total_ms / 5
    out["rank_compute_ms"] = [comm.allreduce(mine if q == rank else 0.0, "sum") for q in range(world)]
    eng.destroy()
    del eng
    single = None
    if world > 1 and rank == 0 and not getattr(args, "no_single", False):
        try:
            H = sp.generators.rmat(scale_log2, ef, seed, fmt=sp.HYB)
            H.set_option("beta0", 1)  # the same operation the ranks perform: y = alpha*A*x, no exchange
            H.set_alpha(ALPHA)
            steps1 = max(3, min(args.steps, 10))
            total_ms, _ = sp.time_rotating([H], steps1, 3, False)
            single = {"ms_per_step": total_ms / steps1, "algorithmic_bytes": int(H.algorithmic_bytes())}
            del H
        except Exception as ex:
            single = {"error": str(ex)}
    comm.barrier()
    out["single"] = single
    return out


def c4_summary(c4, world, peak):
    t = c4["ms_per_step"] * 1e-3
    B = c4["B"]
    d = {"workload": "c4_hyb_row_partitioned", "description": f"row-partitioned hybrid ELL+COO, R-MAT 2^{c4['scale_log2']} x {c4['ef']} (config 4), "
         "x_(k+1) = alpha A x_k, one step = all-gather of x + ELL kernel + COO kernel per piece",
         "ms_per_step": c4["ms_per_step"], "gbs": B / t / 1e9, "gflops": 2.0 * c4["nnz"] / t / 1e9,
         "frac_of_8TBs_nominal_per_gpu": B / t / 1e9 / world / NOMINAL_HBM_GBS, "frac_of_measured_peak_per_gpu": B / t / 1e9 / world / peak,
         "algorithmic_bytes": int(B), "nonzeros": c4["nnz"],
         "partition": f"balanced cost, row = its entries + {c4['row_weight']:g} (spmvb200_partition_rows_weighted; 0 = balanced non-zeros)",
         "row_starts": c4["starts"], "per_rank": [{"matrix_size": p[0], "num_coo_entries": p[1], "ell_row_length": p[2]} for p in c4["per_rank"]],
         "exchange": "allgather", "transport": c4["transport"], "recv_bytes_per_step_per_rank": c4["recv_bytes"],
         "overlap": "column split: the entries that reference the rank's own slice of x run during the all-gather" if c4["split"] else "none",
         "rank0_pieces": c4["pieces"], "gpu_launches": c4["gpu_launches"], "x_norm": c4["x_norm"], "parity": c4["parity"],
         "e2e_ms_per_step": c4["e2e_ms_per_step"], "rank_compute_ms_without_exchange": c4.get("rank_compute_ms")}
    single = c4.get("single")
    if single and "ms_per_step" in single:
        d["single_gpu"] = {"ms_per_step": single["ms_per_step"], "gbs": single["algorithmic_bytes"] / (single["ms_per_step"] * 1e-3) / 1e9,
                           "speedup": single["ms_per_step"] / c4["ms_per_step"]}
    elif single:
        d["single_gpu"] = single
    return d


def run_multi_gpu(args):
    sp, D, dist, rank, world, local_rank, comm = _init_ranks()
    peak, peak_src = measured_peak()
    sampler = ClockSampler(local_rank)
    sampler.start()
    rc = 0
    headline_c4 = getattr(args, "workload", None) == "c4_hyb"
    line = None
    if not headline_c4:
        m = measure_c5_partitioned(args, sp, D, comm, rank, world, sampler)
        results, B, N, nnz, n = m["results"], m["B"], m["N"], m["nnz"], m["grid"]
        timed = {k: v for k, v in results.items() if "ms_per_step" in v}
        failed = [k for k, v in results.items() if not v["parity"]["ok"]]
        # one-GPU time of the SAME matrix and operation measured in the same job on rank 0 (strong-scaling reference)
        single = None
        if world > 1 and rank == 0 and not getattr(args, "no_single", False):
            try:
                full = sp.generators.stencil(sp.STENCIL_3D27, n, n, n, fmt=sp.CSR)
                full.set_option("beta0", 1)
                full.set_alpha(m["scale"])
                k1 = max(3, min(args.steps, 10))
                total_ms, _ = sp.time_rotating([full], k1, 3, False)
                single = total_ms / k1
                del full
            except Exception as ex:  # e.g. not enough memory left
                results["single_gpu_error"] = str(ex)
        comm.barrier()
        if failed or not timed:
            rc = 3
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "parity check failed", "n_gpus": world,
                                  "parity": {k: results[k]["parity"] for k in failed}}), flush=True)
        else:
            best = min(timed, key=lambda k: timed[k]["ms_per_step"])
            r = timed[best]
            t = r["ms_per_step"] * 1e-3
            per_rank_bytes = B / world
            line = {
                "metric": METRIC, "value": r["gbs"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "c5_csr", "description": f"row-partitioned CSR, 3D 27-point {n}^3 (config 5), "
                           f"x_(k+1) = A x_k / 52, one step = exchange of x + SpMV", "rows": N, "nonzeros": nnz,
                           "algorithmic_bytes": B, "partition": "reference rule ceil(rows/P) (csr-matrix.cpp:77-83)",
                           "exchange": best, "executor": "spmvb200_dist_* (C ABI; transports: NCCL loaded at run time, or copy-engine "
                                                         "pulls over NVLink out of IPC-mapped peer buffers = the *_peer variants)",
                           "overlap": "interior rows run during the exchange; boundary rows on their own stream as soon as the halo arrives",
                           "l2": "working set per rank far larger than L2"},
                "gflops": r["gflops"], "frac_of_8TBs_nominal_per_gpu": r["gbs"] / world / NOMINAL_HBM_GBS,
                "parity": r["parity"],
                "roofline": {"bound": "hbm", "achieved": per_rank_bytes / t / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": per_rank_bytes / t / 1e9 / peak, "traffic": None,
                             "kernel": r.get("kernel", "csr kernel") + " (per rank; step time includes the exchange)",
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": int(per_rank_bytes)},
                "e2e": {"value": B / (r["e2e_ms_per_step"] * 1e-3) / 1e9, "unit": UNIT,
                        "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N, "ms_per_step": r["e2e_ms_per_step"],
                        "steps": r["e2e_steps"], "max_abs_y": r["e2e_max_abs_y"],
                        "call": "spmvb200_dist_run_host per rank: pinned x slice -> device, exchange + SpMV, y slice -> pinned host, "
                                "upload of step i+1 and download of step i-1 overlapped with step i"},
                "gpu_launches": r["gpu_launches"], "clocks": sampler.summary(*m["marks"][best]), "exchange_variants": results,
            }
            if single:
                line["single_gpu"] = {"ms_per_step": single, "gbs": B / (single * 1e-3) / 1e9, "speedup": single / r["ms_per_step"]}
    if rc == 0 and (headline_c4 or (world == 2 and not args.no_extra)):
        c4 = measure_c4_partitioned(args, sp, D, comm, rank, world)
        if not c4["parity"]["ok"]:
            rc = 3
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "parity check failed (c4_hyb)", "n_gpus": world, "parity": c4["parity"]}), flush=True)
        elif headline_c4:
            summ = c4_summary(c4, world, peak)
            t = c4["ms_per_step"] * 1e-3
            line = {
                "metric": METRIC, "value": summ["gbs"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": c4["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {k: summ[k] for k in ("workload", "description", "nonzeros", "algorithmic_bytes", "partition", "row_starts",
                                                "per_rank", "exchange", "recv_bytes_per_step_per_rank", "overlap", "rank0_pieces")},
                "gflops": summ["gflops"], "frac_of_8TBs_nominal_per_gpu": summ["frac_of_8TBs_nominal_per_gpu"], "parity": c4["parity"],
                "roofline": {"bound": "hbm", "achieved": c4["B"] / world / t / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": c4["B"] / world / t / 1e9 / peak, "traffic": None,
                             "kernel": "ell_kernel+coo kernel (per rank; step time includes the exchange)", "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": int(c4["B"] / world)},
                "e2e": {"value": c4["B"] / (c4["e2e_ms_per_step"] * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8 * c4["N"],
                        "d2h_bytes_per_step": 8 * c4["N"], "ms_per_step": c4["e2e_ms_per_step"],
                        "call": "spmvb200_dist_run_host per rank"},
                "gpu_launches": c4["gpu_launches"], "clocks": sampler.summary(*c4["marks"]), "x_norm": c4["x_norm"],
            }
            if "single_gpu" in summ:
                line["single_gpu"] = summ["single_gpu"]
        elif line is not None:
            line["c4_hyb"] = c4_summary(c4, world, peak)
    sampler.stop()
    if rank == 0 and line is not None and rc == 0:
        print(json.dumps(line), flush=True)
    comm.destroy()
    if world > 1:
        dist.destroy_process_group()
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-extra", action="store_true", help="skip the targets / per-format / configs[3] blocks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline")
    ap.add_argument("--exchange", default=None, choices=["allgather", "halo", "auto"],
                    help="N>1: how x is exchanged (default: measure allgather and halo, headline = faster)")
    ap.add_argument("--no-single", action="store_true", help="N>1: skip the one-GPU run of the same matrix")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        return run_multi_gpu(args)
    return run_single_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
