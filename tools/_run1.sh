set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
tail -5 gpurun_out/s2_pytest.log
L=gpurun_out/s2_sweep_a.log; : > $L
for w in c1_coo c3s_coo c3_coo; do
  python tools/run_workload.py $w --steps 50 --sweep coo.algo=1 >> $L 2>&1
  python tools/run_workload.py $w --steps 50 --sweep coo.algo=2 --sweep coo.items=2,4,8 --sweep coo.threads=128,256 >> $L 2>&1
done
python tools/run_workload.py c3_coo_atomic --steps 20 --sweep coo.algo=0,2 >> $L 2>&1
python tools/run_workload.py c4s_hyb --steps 50 --sweep coo.algo=1,2 >> $L 2>&1
for w in c1_csr c1_ell c1_hyb c2_csr c2_ell c5s_csr; do
  python tools/run_workload.py $w --steps 200 --sweep independent_launches=0,-1 >> $L 2>&1
done
cat $L
