python -m pytest tests -m gpu -q > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
python bench.py --steps 2000 --warmup 10 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo BENCH_EXIT=$?
python bench.py --steps 20 --warmup 3 --no-extra --no-cpu > gpurun_out/s4_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s4_launches.csv python bench.py --steps 20 --warmup 3 --no-extra --no-cpu > gpurun_out/s4_ncu_launch.log 2>&1
tail -c 300 gpurun_out/s4_bench.err
