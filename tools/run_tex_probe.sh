# texture-path gather experiment (coo.xload = 5 / 6) + host-link yardstick, one GPU
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gather_path or column_blocked" > gpurun_out/u1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u1_pytest.log
python tools/run_workload.py c3_coo --steps 20 --opt coo.hot=-1 --sweep coo.xload=0,5,6,0 > gpurun_out/u1_sweep_tex.log 2>&1
python tools/run_workload.py c3_coo --steps 20 --opt coo.hot=-1 --opt coo.xload=5 --sweep coo.carveout=0,50,100 >> gpurun_out/u1_sweep_tex.log 2>&1
python tools/run_workload.py c4_hyb --steps 10 --opt coo.hot=-1 --sweep coo.xload=0,5,6 >> gpurun_out/u1_sweep_tex.log 2>&1
cat gpurun_out/u1_sweep_tex.log
python tools/pcie_yardstick.py > gpurun_out/u1_pcie.json 2> gpurun_out/u1_pcie.err; cat gpurun_out/u1_pcie.json
python tools/pcie_yardstick.py --mb 256 >> gpurun_out/u1_pcie.json 2>> gpurun_out/u1_pcie.err; tail -1 gpurun_out/u1_pcie.json
