python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; tail -3 gpurun_out/s3_pytest.log
python bench.py --steps 2000 --warmup 10 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo BENCH_EXIT=$?
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/s3_ref.json 2>> gpurun_out/s3_bench.err
for w in c1_csr c2_ell c5s_csr; do
k=csr_flat; [ $w = c2_ell ] && k=ell_kernel
python tools/run_workload.py $w --steps 3 --warmup 1 > gpurun_out/s3_plain_$w.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/s3_prof_$w python tools/run_workload.py $w --steps 3 --warmup 1 > gpurun_out/s3_ncu_$w.log 2>&1
done
tail -c 400 gpurun_out/s3_bench.err
