python -m pytest tests -m gpu -q > gpurun_out/s8_pytest.log 2>&1; tail -3 gpurun_out/s8_pytest.log
L=gpurun_out/s8_sweep.log; : > $L
for k in 0 20 21; do echo "[coo.col_block_log2=$k]" >> $L; python tools/run_workload.py c4_hyb --steps 10 --gopt coo.col_block_log2=$k >> $L 2>&1; done
cat $L
python tools/xgather_report.py --workloads c1_csr --layout-experiment --out gpurun_out/s8_layout.json > gpurun_out/s8_layout.log 2>&1; tail -4 gpurun_out/s8_layout.log
python tools/run_workload.py c4_hyb --steps 3 --warmup 1 > gpurun_out/s8_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:coo_warp4 -s 2 -c 1 -o gpurun_out/s8_prof_c4_hyb_blocked python tools/run_workload.py c4_hyb --steps 3 --warmup 1 > gpurun_out/s8_ncu.log 2>&1
