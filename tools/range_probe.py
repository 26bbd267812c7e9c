#!/usr/bin/env python
"""Measure DRAM traffic IN THE TIMED CONDITION: N launches rotating over the same L2-cold copies bench.py uses, between
cudaProfilerStart / cudaProfilerStop, for ncu's range replay without cache control:

    ncu --replay-mode app-range --cache-control none --clock-control none --profile-from-start off \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/range.csv \
        python tools/range_probe.py c2_ell --launches 200

The per-kernel ncu captures flush the caches and serialise the launches, so their dram__bytes and durations describe an
isolated cold launch; this capture sums the traffic of the whole pipelined sequence, so bytes / launches is the traffic
of a launch as bench.py times it.  Prints the same sequence's CUDA-event time per launch when run without ncu.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import spmv_cache_trace_b200 as sp  # noqa: E402
from bench import free_device_bytes, l2_cold_copies, make_workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--launches", type=int, default=200)
    ap.add_argument("--copies", type=int, default=0)
    args = ap.parse_args()
    cudart = ctypes.CDLL("libcudart.so")
    make, desc = make_workload(sp, args.workload)
    A = make()
    copies = args.copies or l2_cold_copies(A, sp.device_props(0)["l2_bytes"], free_device_bytes())
    mats = [A] + [make() for _ in range(copies - 1)]
    for m in mats:
        m.prepare()
    total_ms, _ = sp.time_rotating(mats, args.launches, 50, False)  # warm-up + the event-timed figure
    for m in mats:
        m.sync()
    assert cudart.cudaProfilerStart() == 0
    for k in range(args.launches):
        mats[k % copies].spmv()
    for m in mats:
        m.sync()
    assert cudart.cudaProfilerStop() == 0
    inf = A.info
    print(json.dumps({"workload": args.workload, "description": desc, "kernel": A.kernel_name, "launches": args.launches,
                      "copies": copies, "algorithmic_bytes": A.algorithmic_bytes(), "event_ms_per_launch": total_ms / args.launches,
                      "xy_bytes_in_cycle": int(copies * (inf.x_size + inf.y_size))}))


if __name__ == "__main__":
    main()
