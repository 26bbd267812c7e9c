python tools/run_workload.py c5s_ell --steps 20 --sweep ell.rows_per_thread=1,2,4 --sweep ell.block=128,256 2>&1 | tail -6
python tools/run_workload.py c5s_csr --steps 20 2>&1 | tail -1
