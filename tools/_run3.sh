python -m pytest tests -m gpu -x -q -k "rows_cut or rmat or poisson2D or config1" 2>&1 | tail -3
L=gpurun_out/s2_sweep_b.log; : > $L
for w in c1_coo c3s_coo c3_coo c3_coo_atomic c4s_hyb; do
  python tools/run_workload.py $w --steps 50 --sweep coo.algo=2,4 --sweep coo.threads=128,256 >> $L 2>&1
done
python tools/run_workload.py c4_hyb --steps 10 --sweep coo.algo=2,4 >> $L 2>&1
cat $L
