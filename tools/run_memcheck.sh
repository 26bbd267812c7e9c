# compute-sanitizer memcheck over the kernels that changed late in round 2, on small cases (one tool, one call)
T='tests/test_gpu_parity.py tests/test_gpu_dist.py'
K='sliced_csr_index_runs or gather_path or drop_and_rebuild or fused_halo_push or (stencil_iteration and 2-halo) or csr_spmv_host_zero_copy'
python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/mc_plain.log 2>&1 && \
compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/mc_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/mc_plain.log; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/mc_memcheck.log | tail -8
