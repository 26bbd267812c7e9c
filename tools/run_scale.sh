# usage: bash tools/run_scale.sh "2 4 8" [grid]     (the driver's scaling run, reproduced by hand)
set -x
NS=${1:-"4 8"}
export SPMV_BENCH_GRID=${2:-512}
for N in $NS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s_n$N.json 2> gpurun_out/s_n$N.err
  echo "N=$N rc=$?"
  tail -c 1500 gpurun_out/s_n$N.err
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/s_n$N.json").read().strip().splitlines()[-1])
    print({k: l[k] for k in ("n_gpus", "ms_per_step", "value")}, l["config"]["exchange"], l.get("single_gpu"), l["e2e"]["ms_per_step"])
    for m, v in l["exchange_variants"].items():
        print(m, v.get("ms_per_step"), v["parity"]["ok"], v.get("e2e_ms_per_step"))
    if "c4_hyb" in l:
        c = l["c4_hyb"]; print("c4_hyb", c["ms_per_step"], c["transport"], c.get("single_gpu"), c["parity"]["ok"], c["e2e_ms_per_step"])
except Exception as e:
    print("no line:", e)
PY
done
