# end-to-end (host buffers) step of config 5: kernel-read y_old (host.zero_copy 1) vs DMA-uploaded y_old (2), by chunk count
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spmv_host" > gpurun_out/u2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u2_pytest.log
python - > gpurun_out/u2_e2e.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import spmv_cache_trace_b200 as sp
A = sp.generators.stencil(sp.STENCIL_3D27, 512, 512, 512)
n = A.rows
xb, yb = sp.PinnedBuffer(n), sp.PinnedBuffer(n)
xb.array[:] = 1.0; yb.array[:] = 0.0
for zc, chunks in ((1, 16), (3, 16), (3, 32), (3, 8), (2, 32), (3, 16), (1, 16)):
    A.set_option("host.zero_copy", zc)
    A.set_option("host.chunks", chunks)
    ms = sp.time_host_rotating([A], [xb.array], [yb.array], 8, 2)
    print("host.zero_copy", zc, "host.chunks", chunks, "ms/step %.3f" % (ms / 8), "GB/s %.1f" % (A.algorithmic_bytes() / (ms / 8) / 1e6), flush=True)
PY
cat gpurun_out/u2_e2e.log
