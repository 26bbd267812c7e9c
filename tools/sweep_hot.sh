python tools/run_workload.py c3_coo --steps 20 --opt coo.hot=-1 --sweep coo.xload=0,4 > gpurun_out/t6_sweep_cpasync.log 2>&1
python tools/run_workload.py c3_coo --steps 20 --opt coo.hot=-1 --opt coo.xload=4 --sweep coo.carveout=25,50,75,100 >> gpurun_out/t6_sweep_cpasync.log 2>&1
cat gpurun_out/t6_sweep_cpasync.log
