python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q -k "read_modify_write or accumulates or alpha_and_beta0 or ordering or one_plan or config5" 2>&1 | tail -4
python tools/run_workload.py c5_csr --steps 20 --warmup 3 --sweep csr.rmw=-1,1,0 2>&1 | tail -4
python tools/run_workload.py c5s_csr --steps 50 --warmup 5 --sweep csr.rmw=-1,1 2>&1 | tail -3
