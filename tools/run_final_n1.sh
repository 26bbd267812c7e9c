# round-end verification on one GPU: the driver's three steps (GPU tests, smoke, bench) + the reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/v1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/v1_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/v1_bench.json 2> gpurun_out/v1_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/v1_ref.json 2> gpurun_out/v1_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/v1_bench.json", "gpurun_out/v1_ref.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d.get("value"), d.get("ms_per_step"), d.get("e2e"), d.get("parity", {}).get("ok"))
PY
