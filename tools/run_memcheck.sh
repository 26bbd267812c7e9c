# compute-sanitizer memcheck over the kernels that are new in round 2, on small cases (one tool, one call)
T='tests/test_gpu_parity.py tests/test_gpu_dist.py'
K='hot_column or ping_pong or spmv_host_zero_copy or beta0_on or weighted or (stencil_iteration and 2-halo) or (column_split and 0) or (run_host and 2-halo)'
python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/mc_plain.log 2>&1 && \
compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/mc_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/mc_plain.log; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/mc_memcheck.log | tail -8
python tools/run_workload.py c1_ell --copies 17 --steps 2000 --warmup 200 --sweep ell.rows_per_thread=1,2,4 --sweep ell.block=64,128,256 > gpurun_out/sweep_t_ell_cold.log 2>&1
python tools/run_workload.py c2_ell --copies 8 --steps 2000 --warmup 200 --sweep ell.rows_per_thread=1,2,4 --sweep ell.block=64,128,256 >> gpurun_out/sweep_t_ell_cold.log 2>&1
python tools/run_workload.py c1_csr --copies 17 --steps 2000 --warmup 200 --sweep csr.threads=64,128,256 --sweep csr.entries=4,8 >> gpurun_out/sweep_t_ell_cold.log 2>&1
cat gpurun_out/sweep_t_ell_cold.log
