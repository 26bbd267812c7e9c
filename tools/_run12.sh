L=gpurun_out/s10_sweep_sliced.log; : > $L
python tools/run_workload.py c5s_csr --steps 50 --sweep csr.algo=5 --sweep csr.batch=2,4,8 >> $L 2>&1
python tools/run_workload.py c5_csr --steps 10 --sweep csr.algo=5 >> $L 2>&1
cat $L
python tools/run_workload.py c5s_csr --steps 3 --warmup 1 > gpurun_out/s10_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:csr_sliced -s 2 -c 1 -o gpurun_out/s10_prof_c5s_sliced python tools/run_workload.py c5s_csr --steps 3 --warmup 1 > gpurun_out/s10_ncu.log 2>&1
python tools/run_workload.py c5s_ell --steps 3 --warmup 1 > gpurun_out/s10_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ell_kernel -s 2 -c 1 -o gpurun_out/s10_prof_c5s_ell python tools/run_workload.py c5s_ell --steps 3 --warmup 1 > gpurun_out/s10_ncu2.log 2>&1
