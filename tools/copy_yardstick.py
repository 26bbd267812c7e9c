import sys, numpy as np
sys.path.insert(0, ".")
import spmv_cache_trace_b200 as sp
for nbytes in (40_000_000, 105_000_000):
    ms = sp.time_copy(nbytes, copies=8, reps=300, warmup=30)
    print("copy", nbytes, "us median/min", float(np.median(ms))*1e3, float(ms.min())*1e3, "frac of 8TB/s", 2*nbytes/(np.median(ms)*1e-3)/1e9/8000)
