# short-row matrices on the sliced kernel + ncu evidence for the final headline kernel
python tools/run_workload.py c1_csr --copies 17 --steps 2000 --warmup 200 --sweep csr.algo=0,5,0,5 > gpurun_out/w6_sweep.log 2>&1
python tools/run_workload.py c1_csr --copies 17 --steps 2000 --warmup 200 --opt csr.algo=5 --sweep csr.batch=2,8 --sweep csr.threads=128 >> gpurun_out/w6_sweep.log 2>&1
python tools/run_workload.py c1_csr --copies 17 --steps 2000 --warmup 200 --opt csr.algo=5 --sweep csr.threads=256,512 >> gpurun_out/w6_sweep.log 2>&1
python tools/run_workload.py c2_csr --copies 8 --steps 2000 --warmup 200 --opt csr.algo=5 --sweep csr.batch=2,4,8 >> gpurun_out/w6_sweep.log 2>&1
cat gpurun_out/w6_sweep.log
W="python tools/run_workload.py c5_csr --steps 2 --warmup 1"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:csr_sliced -s 1 -c 2 --csv \
    --log-file gpurun_out/r03_traffic_c5_csr.csv $W > gpurun_out/w6_ncu_c5.log 2>&1
echo "c5 traffic rc=$?"
W2="python tools/run_workload.py c5s_csr --steps 2 --warmup 1"
$W2 > gpurun_out/w6_plain_c5s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:csr_sliced -s 1 -c 1 -o gpurun_out/r03_prof_c5s_csr_runs -f $W2 > gpurun_out/w6_ncu_c5s.log 2>&1
echo "c5s full rc=$?"; tail -2 gpurun_out/w6_ncu_c5s.log
