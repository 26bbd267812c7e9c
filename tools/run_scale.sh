# N = 4 and N = 8 on one 8-GPU box (the driver's scaling run, reproduced by hand)
set -x
for N in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/t7_n$N.json 2> gpurun_out/t7_n$N.err
  echo "N=$N rc=$?"
  tail -c 600 gpurun_out/t7_n$N.err
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/t7_n$N.json").read().strip().splitlines()[-1])
    print({k: l[k] for k in ("n_gpus", "ms_per_step", "value")}, l.get("single_gpu"), l["e2e"]["ms_per_step"])
    for m, v in l["exchange_variants"].items():
        print(m, v["ms_per_step"], v["parity"]["ok"], v["e2e_ms_per_step"])
except Exception as e:
    print("no line:", e)
PY
done
