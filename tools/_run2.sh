set -x
for w in c3_coo c1_coo c4s_hyb; do
python tools/run_workload.py $w --steps 3 --warmup 1 > gpurun_out/s2_plain_$w.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:coo_warp -s 2 -c 1 -o gpurun_out/s2_prof_$w python tools/run_workload.py $w --steps 3 --warmup 1 > gpurun_out/s2_ncu_$w.log 2>&1
done
ls -la gpurun_out | tail -8
