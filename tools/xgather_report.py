#!/usr/bin/env python
"""Predicted x-gather miss traffic (the reference's cache model, spmvb200_cache_trace) next to the DRAM
bytes ncu measured for the same kernels -- north-star item (4), SURVEY 8(f)1.

    python tools/xgather_report.py [--out profiles/r01_xgather_model.json] [--workloads c1_csr,c3_coo,...]

The matrices are the device-generated bench workloads (their index arrays are copied back; the model
itself is host code).  Model: fully associative LRU of the size of this GPU's L2, 32 B lines (the DRAM
sector), one cache per part; "bypass" = the matrix streams do not allocate (the kernels' evict-first
policy), "lru" = the reference's plain LRU for every reference.  Measured bytes come from the committed
ncu captures (profiles/ncu_traffic.json, key "<workload>": {"read": .., "write": ..}).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import spmv_cache_trace_b200 as sp  # noqa: E402
from tools.run_workload import factory  # noqa: E402


def total(parts, keys):
    return int(sum(p[k] for p in parts for k in keys))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_xgather_model.json"))
    ap.add_argument("--workloads", default="c1_csr,c1_coo,c2_ell,c3s_coo,c3_coo,c5s_csr")
    ap.add_argument("--line", type=int, default=32)
    ap.add_argument("--layout-experiment", action="store_true")
    args = ap.parse_args()
    l2 = sp.device_props(0)["l2_bytes"]
    try:
        measured = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        measured = {}
    doc = {"cache_bytes": int(l2), "line_bytes": args.line, "model": "fully associative LRU, spmvb200_cache_trace", "workloads": []}
    for name in args.workloads.split(","):
        A = factory(name)()
        inf = A.info
        entry = {"workload": name, "rows": int(inf.rows), "nonzeros": int(inf.num_entries),
                 "algorithmic_bytes": int(A.algorithmic_bytes()), "x_bytes": int(8 * inf.columns)}
        for label, bypass in (("bypass", True), ("lru", False)):
            t0 = time.time()
            parts = sp.cache_model.matrix(A, l2, args.line, parts=1, shared=False, stream_bypass=bypass)
            x_miss = total(parts, ("misses_x_local", "misses_x_remote")) * args.line
            every = total(parts, ("misses_index", "misses_column_index", "misses_value", "misses_x_local", "misses_x_remote",
                                  "misses_y_local", "misses_y_remote")) * args.line
            entry[label] = {"x_gather_miss_bytes": x_miss, "x_gather_refetch_factor": x_miss / max(1, 8 * inf.columns),
                            "predicted_dram_read_bytes": every, "x_references": total(parts, ("x_references",)),
                            "model_seconds": round(time.time() - t0, 2)}
        m = measured.get(name)
        if isinstance(m, dict):
            entry["ncu"] = m
            entry["ncu_read_over_predicted"] = m["read"] / max(1, entry["bypass"]["predicted_dram_read_bytes"])
        doc["workloads"].append(entry)
        print(json.dumps(entry), flush=True)
        if name == "c5s_csr":  # the multi-GPU partition: 8 ranks, balanced non-zeros, what each must receive
            starts = sp.partition.rows_nnz(A, 8)
            parts = sp.cache_model.matrix(A, l2, args.line, parts=8, starts=starts, shared=False, stream_bypass=True)
            doc["partition_c5s_8"] = {
                "starts": [int(s) for s in starts],
                "x_remote_miss_bytes": [p["misses_x_remote"] * args.line for p in parts],
                "x_local_miss_bytes": [p["misses_x_local"] * args.line for p in parts],
                "x_remote_references": [p["x_remote_references"] for p in parts],
                "halo_plane_bytes": 256 * 256 * 8,
                "note": "remote x misses of a rank = the x elements it must receive: one 256x256 plane per neighbour "
                        "(halo exchange) instead of the 7/8 of x an all-gather delivers",
            }
            print(json.dumps(doc["partition_c5s_8"]), flush=True)
        del A
    if args.layout_experiment:
        # Config 4 at 1/16 scale (R-MAT 2^22 x 32, hybrid) with 1/16 of the L2: the row-sorted COO tail against the
        # column-blocked one.  The model walks the entries in the order the kernel does.
        cache = l2 // 16
        exp = {"matrix": "R-MAT 2^22 x 32, hybrid (config 4 at 1/16 scale)", "cache_bytes": int(cache), "line_bytes": args.line,
               "x_bytes": 8 << 22, "layouts": []}
        for label, k in (("row-sorted tail", -1), ("column blocks of 2^18 (x block = cache/4)", 18),
                         ("column blocks of 2^19 (x block = cache/2)", 19)):
            sp.set_global_option("coo.col_block_log2", k)
            try:
                H = sp.generators.rmat(22, 32, 0x5EED0004, fmt=sp.HYB)
            finally:
                sp.set_global_option("coo.col_block_log2", 0)
            r = sp.cache_model.matrix(H, cache, args.line, parts=1, shared=False, stream_bypass=True)[0]
            exp["layouts"].append({"layout": label, "applied_log2": H.get_option("coo.col_block_log2"),
                                   "x_gather_miss_bytes": (r["misses_x_local"] + r["misses_x_remote"]) * args.line,
                                   "y_miss_bytes": (r["misses_y_local"] + r["misses_y_remote"]) * args.line,
                                   "stream_bytes": (r["misses_index"] + r["misses_column_index"] + r["misses_value"]) * args.line})
            print(json.dumps(exp["layouts"][-1]), flush=True)
            del H
        doc["layout_experiment"] = exp
    with open(args.out, "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
