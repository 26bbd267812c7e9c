python -m pytest tests -m gpu -x -q -k "rows_cut or 64bit or config or stencil27 or row_partition or poisson2D or ref_vectors or degenerate" 2>&1 | tail -3
L=gpurun_out/s9_sweep_sliced.log; : > $L
for w in c5s_csr c2_csr c1_csr; do
  python tools/run_workload.py $w --steps 50 --sweep csr.algo=4 >> $L 2>&1
  python tools/run_workload.py $w --steps 50 --sweep csr.algo=5 --sweep csr.batch=2,4,8 >> $L 2>&1
done
python tools/run_workload.py c5_csr --steps 10 --sweep csr.algo=4,5 >> $L 2>&1
cat $L
