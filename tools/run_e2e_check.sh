python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spmv_host or hot_column" > gpurun_out/t11_pytest.log 2>&1; tail -5 gpurun_out/t11_pytest.log
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import spmv_cache_trace_b200 as sp
A = sp.generators.stencil(sp.STENCIL_3D27, 512, 512, 512)
n = A.rows
xb, yb = sp.PinnedBuffer(n), sp.PinnedBuffer(n)
xb.array[:] = 1.0; yb.array[:] = 0.0
for chunks in (1, 4, 8, 16, 32, 64):
    A.set_option("host.chunks", chunks)
    ms = sp.time_host_rotating([A], [xb.array], [yb.array], 5, 1)
    print("host.chunks", chunks, "ms/step", ms / 5, "GB/s", A.algorithmic_bytes() / (ms / 5) / 1e6, flush=True)
PY
