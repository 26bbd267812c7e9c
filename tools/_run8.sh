python -m pytest tests -m gpu -x -q -k "column_blocked or rows_cut or rmat or hybrid" 2>&1 | tail -3
L=gpurun_out/s7_sweep_blocks.log; : > $L
python tools/run_workload.py c4_hyb --steps 10 --gopt coo.col_block_log2=-1 >> $L 2>&1
for k in 0 22 23 24; do python tools/run_workload.py c4_hyb --steps 10 --gopt coo.col_block_log2=$k >> $L 2>&1; done
for k in -1 21 22 23; do python tools/run_workload.py c3_coo --steps 20 --gopt coo.col_block_log2=$k >> $L 2>&1; done
cat $L
