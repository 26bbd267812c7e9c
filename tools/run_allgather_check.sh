for N in ${1:-8 4}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 --exchange allgather --no-extra > gpurun_out/ag_n$N.json 2> gpurun_out/ag_n$N.err
echo "N=$N rc=$?"; tail -c 400 gpurun_out/ag_n$N.err
python - <<PY
import json
l = json.loads(open("gpurun_out/ag_n$N.json").read().strip().splitlines()[-1])
print(l.get("single_gpu"))
for m, v in l["exchange_variants"].items():
    print(m, v.get("ms_per_step"), v["parity"]["ok"], v.get("e2e_ms_per_step"))
PY
done
