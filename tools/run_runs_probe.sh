# diagonal slices of the sliced CSR kernel: tests, then timings
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q -x -k "sliced or spmv_host or ping_pong or alpha or push or stencil_iteration" > gpurun_out/w5_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/w5_pytest.log
python tools/run_workload.py c5s_csr --steps 20 --sweep csr.index_runs=-1,0 > gpurun_out/w5_sweep.log 2>&1
python tools/run_workload.py c5_csr --steps 10 --sweep csr.index_runs=0,0 >> gpurun_out/w5_sweep.log 2>&1
python tools/run_workload.py c2_csr --copies 8 --steps 2000 --warmup 200 --sweep csr.algo=0,5 >> gpurun_out/w5_sweep.log 2>&1
cat gpurun_out/w5_sweep.log
