#!/usr/bin/env python
"""Generates tests/golden/cache_trace.json with the UNMODIFIED reference cache simulator
(oracle/_ref/libspmvref.so: replacement::LRU + trace_cache_misses over the reference's own
memory reference strings).  Run in the build container, where /root/reference exists:

    python tests/golden/make_cache_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.oracle import Ref  # noqa: E402
from test_cache_model import random_case, ref_matrix_local_remote  # noqa: E402


def main():
    cases = []
    for seed, rows, cols, T, cache, line, warm in ((3, 500, 500, 2, 8192, 64, False), (4, 400, 700, 3, 4096, 64, True),
                                                   (5, 700, 400, 1, 16384, 128, False), (6, 600, 600, 4, 32768, 32, True)):
        i, j, a = random_case(seed, rows, cols)
        R = Ref().from_entries(rows, cols, i, j, a)
        R.convert("csr")
        csr = ref_matrix_local_remote(R.cache_trace(T, cache, line, warm), T)
        R.convert("coo-atomic")
        coo = ref_matrix_local_remote(R.cache_trace(T, cache, line, warm), T)
        cases.append(dict(seed=seed, rows=rows, cols=cols, threads=T, cache_bytes=cache, line_bytes=line, warmup=warm,
                          csr=[list(t) for t in csr], coo_atomic=[list(t) for t in coo]))
    with open(os.path.join(HERE, "cache_trace.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_cache_golden.py via oracle/_ref (reference LRU simulator)", "cases": cases}, f, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
