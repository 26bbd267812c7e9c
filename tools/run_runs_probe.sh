# short-row stencils at scale: flat CSR vs sliced CSR with diagonal slices (and ELL for reference)
python tools/run_workload.py big7_csr --steps 20 --sweep csr.algo=0,5 > gpurun_out/w7_sweep.log 2>&1
python tools/run_workload.py big5_csr --steps 20 --sweep csr.algo=0,5 >> gpurun_out/w7_sweep.log 2>&1
python tools/run_workload.py big7_ell --steps 20 >> gpurun_out/w7_sweep.log 2>&1
cat gpurun_out/w7_sweep.log
