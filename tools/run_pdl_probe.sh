nvidia-smi topo -m > gpurun_out/t8_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/t8_topo.txt; cat gpurun_out/t8_topo.txt
X=spmv_cache_trace_b200/lib/libspmvb200_xcoherent.so
for pdl in 2 0; do
  timeout 300 python tools/pdl_probe.py --pdl $pdl --devices 0,1 >> gpurun_out/t8_pdl_probe.jsonl 2>> gpurun_out/t8_pdl_probe.err
  SPMVB200_LIB=$PWD/$X timeout 300 python tools/pdl_probe.py --pdl $pdl --devices 0,1 >> gpurun_out/t8_pdl_probe.jsonl 2>> gpurun_out/t8_pdl_probe.err
done
timeout 300 python tools/pdl_probe.py --pdl 2 --devices 0,0 >> gpurun_out/t8_pdl_probe.jsonl 2>> gpurun_out/t8_pdl_probe.err
cat gpurun_out/t8_pdl_probe.jsonl; tail -5 gpurun_out/t8_pdl_probe.err
