#!/usr/bin/env python
"""bench.py -- SpMV throughput of the B200-native engine on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--no-extra]

One JSON line on stdout (rank 0).  A "step" is one y += A*x over the whole matrix.

N = 1   headline workload = BASELINE configs[1]: ELLPACK SpMV on the 3D 7-point Poisson matrix of a
        128^3 grid (2 097 152 rows, 14 581 760 nnz, W = 7), fp64.  `value` = effective GB/s
        = (matrix_size + x_size + y_size) / t as the reference prints those sizes, inputs resident in
        HBM, K launches timed with CUDA events; the launches rotate over independent copies of the
        workload so no launch finds its operands in L2 (the working set, 210 MB, is of the order of
        the 126 MB L2).  "formats" carries the same measurement for every format of config 1
        (2D 5-point 1000x1000: CSR, ELL, COO, hybrid) and for configs 3-5 (R-MAT COO 2^24x16,
        R-MAT hybrid 2^26x32, 27-point CSR 512^3), each with its own roofline fraction.
N > 1   BASELINE configs[4]: row-partitioned CSR SpMV on the 27-point stencil of a 512^3 grid
        (134 217 728 rows, 3 609 741 304 nnz), strong scaling, a step = one SpMV plus the exchange of
        x between ranks over NCCL; see spmv_cache_trace_b200/distributed.py.
        `--workload c4_hyb` (N = 2): BASELINE configs[3], the hybrid ELL+COO matrix (R-MAT 2^26 x 32)
        cut into row blocks of equal non-zeros, x all-gathered between iterations.
--impl reference   the reference's own OpenMP kernels (oracle/_ref, compiled from the unmodified
        reference sources) on this box's host cores, same workload, same metric.

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
# workloads (single GPU)
# ------------------------------------------------------------------------------------------------

def make_workload(sp, name: str):
    """Returns (factory() -> DeviceMatrix, description)."""
    g = sp.generators
    table = {
        "c1_csr": (lambda: g.stencil(sp.STENCIL_2D5, 1000, 1000, 1, sp.CSR), "CSR, 2D 5-point Poisson 1000x1000 (config 1)"),
        "c1_ell": (lambda: g.stencil(sp.STENCIL_2D5, 1000, 1000, 1, sp.ELL), "ELL, 2D 5-point Poisson 1000x1000 (config 1 matrix)"),
        "c1_coo": (lambda: g.stencil(sp.STENCIL_2D5, 1000, 1000, 1, sp.COO), "COO, 2D 5-point Poisson 1000x1000 (config 1 matrix)"),
        "c1_hyb": (lambda: g.stencil(sp.STENCIL_2D5, 1000, 1000, 1, sp.HYB), "hybrid, 2D 5-point Poisson 1000x1000 (config 1 matrix)"),
        "c2_ell": (lambda: g.stencil(sp.STENCIL_3D7, 128, 128, 128, sp.ELL), "ELL, 3D 7-point Poisson 128^3 (config 2)"),
        "c2_csr": (lambda: g.stencil(sp.STENCIL_3D7, 128, 128, 128, sp.CSR), "CSR, 3D 7-point Poisson 128^3 (config 2 matrix)"),
        "c3_coo": (lambda: g.rmat(24, 16, 0x5EED0003, fmt=sp.COO), "COO (segmented), R-MAT 2^24 x 16 (config 3)"),
        "c3_coo_atomic": (lambda: g.rmat(24, 16, 0x5EED0003, fmt=sp.COO, coo_mode=sp.COO_ATOMIC), "COO (atomic), R-MAT 2^24 x 16 (config 3 matrix)"),
        "c4_hyb": (lambda: g.rmat(26, 32, 0x5EED0004, fmt=sp.HYB), "hybrid, R-MAT 2^26 x 32 (config 4)"),
        "c5_csr": (lambda: g.stencil(sp.STENCIL_3D27, 512, 512, 512, sp.CSR), "CSR, 3D 27-point 512^3 (config 5, one GPU)"),
    }
    return table[name]


def measure_device(sp, name, steps, warmup, l2_bytes, peak, per_launch=True, max_copies=8):
    """Device-resident timing of one workload; returns (result dict, matrices kept alive)."""
    make, desc = make_workload(sp, name)
    A = make()
    B = A.algorithmic_bytes()
    inf = A.info
    # Copies so that consecutive launches never reuse L2 contents: cycle length >= 3 x L2.
    copies = 1 if B >= 3 * l2_bytes else min(max_copies, int(math.ceil(3.0 * l2_bytes / B)))
    mats = [A] + [make() for _ in range(copies - 1)]
    total_ms, per = sp.time_rotating(mats, steps, warmup, per_launch)
    t = total_ms * 1e-3 / steps
    res = {
        "workload": name, "description": desc, "kernel": A.kernel_name,
        "rows": int(inf.rows), "columns": int(inf.columns), "nonzeros": int(inf.num_entries),
        "matrix_size": int(inf.matrix_size), "algorithmic_bytes": int(B), "flops": int(2 * inf.num_entries),
        "ms_per_step": total_ms / steps, "gbs": B / t / 1e9, "gflops": 2.0 * inf.num_entries / t / 1e9,
        "frac_of_8TBs": B / t / 1e9 / NOMINAL_HBM_GBS, "frac_of_measured_peak": B / t / 1e9 / peak,
        "l2_cold_copies": copies,
        "traffic": ncu_traffic(name),  # dram__bytes_read + write of one launch from the committed ncu capture, if any
    }
    if inf.format == sp.HYB:
        res.update(ell_row_length=int(inf.ell_row_length), num_coo_entries=int(inf.num_coo_entries))
    if inf.format == sp.ELL:
        res.update(ell_row_length=int(inf.ell_row_length))
    if per is not None:
        res["kernel_ms_mean"] = float(np.mean(per))
        res["kernel_ms_median"] = float(np.median(per))
        res["kernel_ms_min"] = float(np.min(per))
    res["ordered_launches"] = measure_ordered(sp, mats, steps, warmup, B, peak)
    return res, mats


def measure_ordered(sp, mats, steps, warmup, B, peak):
    """The same K launches with full ordering forced ("independent_launches" = -1: every kernel executes
    griddepcontrol.wait before it touches x or y).  By default the library orders two launches only when
    one writes what the other reads (it tracks the x/y ranges in flight on its own streams); in the
    reference protocol -- x constant, y accumulated with reductions -- that is never the case, so the
    drain of one launch overlaps the ramp of the next.  Reported next to the default for comparison."""
    for m in mats:
        m.set_option("independent_launches", -1)
    try:
        total_ms, _ = sp.time_rotating(mats, steps, warmup, False)
    finally:
        for m in mats:
            m.set_option("independent_launches", 0)
    t = total_ms * 1e-3 / steps
    return {"ms_per_step": total_ms / steps, "gbs": B / t / 1e9, "frac_of_8TBs": B / t / 1e9 / NOMINAL_HBM_GBS,
            "frac_of_measured_peak": B / t / 1e9 / peak}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own kernels on this box's host cores
# ------------------------------------------------------------------------------------------------

def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(workload: str, reps: int, warmup: int = 1, threads: int = 0):
    """Times oracle/_ref (the unmodified reference library) -- or the oracle port if it is missing --
    with the protocol of profile_kernel_run, on the SAME matrix as the GPU workload."""
    from oracle.generators_ref import stencil_entries
    from oracle.oracle import Oracle, Ref
    T = threads or host_threads()
    kind_fmt = {"c2_ell": (1, (128, 128, 128), "ell"), "c1_csr": (0, (1000, 1000, 1), "csr"),
                "c1_ell": (0, (1000, 1000, 1), "ell"), "c1_coo": (0, (1000, 1000, 1), "coo"),
                "c1_hyb": (0, (1000, 1000, 1), "hybrid"), "c2_csr": (1, (128, 128, 128), "csr"),
                # config 5 exceeds the reference's int32 sizes: the baseline sample is a 512x512x24 slab
                "c5_csr": (2, (512, 512, 24), "csr")}
    kind, dims, fmt = kind_fmt[workload]
    i, j, a = stencil_entries(kind, *dims)
    n = dims[0] * dims[1] * dims[2]
    sample = f"{fmt.upper()} y+=A*x on the {dims[0]}x{dims[1]}x{dims[2]} stencil matrix ({n} rows, {len(i)} nnz), " \
             f"x=1, 1 warm-up + {reps} timed runs, barrier/steady_clock/barrier (profile-kernel.cpp:137-179)"
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
    wl = args.workload or ("c2_ell" if args.gpus == 1 else "c5_csr")
    if wl not in ("c2_ell", "c1_csr", "c1_ell", "c1_coo", "c1_hyb", "c2_csr", "c5_csr"):
        # R-MAT configs exceed what the reference's int32 conversion can hold in host memory here: their CPU sample is
        # the default one of this GPU count
        wl = "c2_ell" if args.gpus == 1 else "c5_csr"
    t0 = time.perf_counter()
    cb = cpu_baseline(wl, reps=max(args.steps, 1), warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_median"], "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "host_threads": cb["cores"]},
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
    wl = args.workload or "c2_ell"
    sampler = ClockSampler(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0] or 0))
    sampler.start()

    # ---- headline: device-resident, L2-cold ------------------------------------------------------
    make, desc = make_workload(sp, wl)
    A = make()
    B = A.algorithmic_bytes()
    inf = A.info
    copies = 1 if B >= 3 * props["l2_bytes"] else min(8, int(math.ceil(3.0 * props["l2_bytes"] / B)))
    mats = [A] + [make() for _ in range(copies - 1)]
    # W untimed warm-up steps (>= 3).  A launch of this workload lasts ~30 us, so a handful of them does not bring the
    # clocks up after the idle time of matrix generation: small workloads get at least 200 (reported as "warmup").
    warm = max(args.warmup, 3 if copies == 1 else 200)
    sp.time_rotating(mats, warm, 0, False)
    launches0 = sp.launch_count()
    sampler.mark("t0"); t0 = time.perf_counter()
    total_ms, _ = sp.time_rotating(mats, args.steps, 0, False)  # the timed region: EXACTLY K steps
    t1 = time.perf_counter(); sampler.mark("t1")
    gpu_launches = sp.launch_count() - launches0
    _, per = sp.time_rotating(mats, args.steps, 0, True)  # same K steps again, one event pair per launch
    per = per if per is not None else np.array([total_ms / args.steps])
    ordered = measure_ordered(sp, mats, args.steps, 3, B, peak)
    t_step = total_ms * 1e-3 / args.steps
    value = B / t_step / 1e9
    # roofline.achieved: algorithmic bytes of one launch / the kernel's average launch duration over the timed
    # region (CUDA events on the launching stream around the K back-to-back launches; nothing else runs there).
    # The per-launch-event pass right after it gives the duration of an ISOLATED launch: the event records
    # between launches keep consecutive kernels from overlapping their ramp and drain, so it is longer.
    k_ms = total_ms / args.steps
    iso_ms = float(np.mean(per))
    achieved = B / (k_ms * 1e-3) / 1e9

    # ---- end to end: host buffers through the C ABI ---------------------------------------------------
    e2e_steps = max(3, min(args.steps, 50))
    xs = [sp.PinnedBuffer(int(inf.columns)) for _ in mats]
    ys = [sp.PinnedBuffer(int(inf.rows)) for _ in mats]
    for xb, yb in zip(xs, ys):
        xb.array[:] = 1.0
        yb.array[:] = 0.0
    e2e_ms = sp.time_host_rotating(mats, [b.array for b in xs], [b.array for b in ys], e2e_steps, 2)
    e2e_t = e2e_ms * 1e-3 / e2e_steps
    # the result of the e2e steps is checked, not just timed: y = (#steps on that copy) * A*1
    ycheck = float(np.abs(ys[0].array).max())
    sampler.stop()
    clocks = sampler.summary(t0, t1)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": warm,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "description": desc, "rows": int(inf.rows), "nonzeros": int(inf.num_entries),
                   "algorithmic_bytes": int(B), "semantics": "y += A*x (reference Kernel::run)",
                   "l2": f"cold: launches rotate over {copies} independent copies of matrix, x and y "
                         f"({copies * B / 1e6:.0f} MB cycle vs {props['l2_bytes'] / 1e6:.0f} MB L2)",
                   "device": props["name"], "sm_count": props["sm_count"]},
        "gflops": 2.0 * inf.num_entries / t_step / 1e9,
        "frac_of_8TBs_nominal": value / NOMINAL_HBM_GBS,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(wl), "kernel": A.kernel_name, "kernel_ms_mean": k_ms,
                     "isolated_launch_ms_mean": iso_ms, "isolated_launch_ms_median": float(np.median(per)),
                     "isolated_launch_gbs": B / (iso_ms * 1e-3) / 1e9, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(B),
                     "how": "CUDA events on the launching stream around the K launches of the timed region (kernel_ms_mean = "
                            "their average duration); isolated_launch_* = one event pair per launch in a second pass",
                     "note": "peak is the driver's COPY bandwidth (read+write); a read-dominated stream can exceed it"},
        "e2e": {"value": B / e2e_t / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(8 * (inf.columns + inf.rows)),
                "d2h_bytes_per_step": int(8 * inf.rows), "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                "call": "spmvb200_spmv_host (pinned host x, y -> device, kernel, y -> host)", "max_abs_y": ycheck},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "ordered_launches": ordered,
    }
    del xs, ys, mats, A

    # ---- CPU baseline (rank 0, bounded sample) ---------------------------------------------------------
    if not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline(wl, reps=10)
        except Exception as e:  # keep the GPU line even if the checker is unavailable
            line["cpu_baseline"] = {"error": repr(e)}

    # ---- the other formats / configurations -----------------------------------------------------------------
    if not args.no_extra:
        extra = []
        names = ["c1_csr", "c1_ell", "c1_coo", "c1_hyb", "c2_csr", "c3_coo", "c3_coo_atomic", "c4_hyb", "c5_csr"]
        for name in names:
            if name == wl:
                continue
            try:
                big = name in ("c3_coo", "c3_coo_atomic", "c4_hyb", "c5_csr")
                # the small matrices take 13-36 us per launch: enough launches that clock ramp-up after the idle
                # time of matrix generation does not colour the result
                r, keep = measure_device(sp, name, 20 if big else 2000, 3 if big else 100, props["l2_bytes"], peak)
                del keep
                extra.append(r)
            except Exception as e:
                extra.append({"workload": name, "error": str(e)})
        line["formats"] = extra
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-extra", action="store_true", help="skip the per-format / per-config block")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline")
    ap.add_argument("--exchange", default=None, choices=["allgather", "halo", "auto"],
                    help="N>1: how x is exchanged (default: measure allgather and halo, headline = faster)")
    ap.add_argument("--no-single", action="store_true", help="N>1: skip the one-GPU run of the same matrix")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from spmv_cache_trace_b200 import distributed
        return distributed.bench_main(args)
    return run_single_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
