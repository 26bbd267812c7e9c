python -m pytest tests -m gpu -x -q -k "rows_cut or rmat or poisson2D or config or 64bit or degenerate" 2>&1 | tail -3
L=gpurun_out/s2_sweep_c.log; : > $L
for w in c1_csr c2_csr c5s_csr c3s_csr c3_csr; do
  python tools/run_workload.py $w --steps 100 --sweep csr.algo=0 >> $L 2>&1
  python tools/run_workload.py $w --steps 100 --sweep csr.algo=4 --sweep csr.threads=64,128,256 >> $L 2>&1
done
cat $L
