python -m pytest tests -m gpu -q > gpurun_out/s11_pytest.log 2>&1; tail -3 gpurun_out/s11_pytest.log
python bench.py --steps 2000 --warmup 10 > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err; echo BENCH_EXIT=$?
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/s11_ref.json 2>> gpurun_out/s11_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s11_smoke.log 2>&1; tail -2 gpurun_out/s11_smoke.log
tail -c 300 gpurun_out/s11_bench.err
