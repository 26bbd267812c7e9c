# ncu evidence for round 2 (one gpurun call; every command runs plain first, then under ncu)
set -x
B="python bench.py --steps 2 --warmup 1 --no-extra --no-cpu"
$B > gpurun_out/n_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c5_csr.csv $B > gpurun_out/n_ncu_bench.log 2>&1
echo "launch list rc=$?"
W="python tools/run_workload.py c5s_csr --steps 2 --warmup 1"
$W > gpurun_out/n_plain_c5s.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csr_sliced -s 1 -c 1 -f -o gpurun_out/r02_prof_c5s_csr_sliced $W > gpurun_out/n_ncu_c5s.log 2>&1
echo "full set rc=$?"
ncu -i gpurun_out/r02_prof_c5s_csr_sliced.ncu-rep --page details > gpurun_out/r02_ncu_c5s_csr_sliced.txt 2>&1
for wl in c1_csr c1_ell c2_ell; do
  R="python tools/range_probe.py $wl --launches 200"
  $R > gpurun_out/r02_range_$wl.json 2> gpurun_out/n_range_$wl.err &&
  ncu --replay-mode app-range --cache-control none --clock-control none --profile-from-start off \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --csv \
      --log-file gpurun_out/r02_range_$wl.csv $R > gpurun_out/n_ncu_range_$wl.log 2>&1
  echo "range $wl rc=$?"; cat gpurun_out/r02_range_$wl.json; tail -4 gpurun_out/r02_range_$wl.csv
done
