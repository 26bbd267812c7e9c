# ncu evidence for round 2, second call: DRAM traffic of the headline kernel at FULL size (single-pass metrics, no replay), and the
# range captures of the rotating L2-cold sequences
set -x
W="python tools/run_workload.py c5_csr --steps 2 --warmup 1"
$W > gpurun_out/n_plain_c5.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:csr_sliced -s 1 -c 2 --csv \
    --log-file gpurun_out/r02_traffic_c5_csr.csv $W > gpurun_out/n_ncu_c5.log 2>&1
echo "c5 traffic rc=$?"; tail -8 gpurun_out/r02_traffic_c5_csr.csv | cut -c150-400
for wl in c1_csr c1_ell c2_ell; do
  R="python tools/range_probe.py $wl --launches 200"
  $R > gpurun_out/r02_range_$wl.json 2> gpurun_out/n_range_$wl.err &&
  ncu --replay-mode app-range --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --csv \
      --log-file gpurun_out/r02_range_$wl.csv $R > gpurun_out/n_ncu_range_$wl.log 2>&1
  echo "range $wl rc=$?"; cat gpurun_out/r02_range_$wl.json; tail -5 gpurun_out/r02_range_$wl.csv | cut -c1-300; tail -3 gpurun_out/n_ncu_range_$wl.log
done
