# the host link with two GPUs copying at once (plain pinned cudaMemcpyAsync, no kernel of ours): is the aggregate shared?
CUDA_VISIBLE_DEVICES=0 python tools/pcie_yardstick.py --mb 512 --reps 8 > gpurun_out/y_gpu0_alone.json 2>/dev/null
CUDA_VISIBLE_DEVICES=0 python tools/pcie_yardstick.py --mb 512 --reps 40 > gpurun_out/y_gpu0_both.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 python tools/pcie_yardstick.py --mb 512 --reps 40 > gpurun_out/y_gpu1_both.json 2>/dev/null &
wait
for f in y_gpu0_alone y_gpu0_both y_gpu1_both; do echo $f; python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print({k: round(v,1) for k,v in d.items() if isinstance(v,float)})"; done
