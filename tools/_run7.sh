for w in c3_coo c4_hyb; do
python tools/run_workload.py $w --steps 3 --warmup 1 > gpurun_out/s5_plain_$w.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:coo_warp4 -s 2 -c 1 -o gpurun_out/s5_prof_$w python tools/run_workload.py $w --steps 3 --warmup 1 > gpurun_out/s5_ncu_$w.log 2>&1
done
python tools/xgather_report.py --out gpurun_out/s5_xgather.json > gpurun_out/s5_xgather.log 2>&1
tail -3 gpurun_out/s5_xgather.log
