# ncu launch list of the bench command (after the same command has exited 0 without ncu)
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/x1_plain.json 2> gpurun_out/x1_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_launches_bench_c5_csr.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/x1_ncu.json 2> gpurun_out/x1_ncu.err; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r03_launches_bench_c5_csr.csv")))
hdr = [r for r in rows if "Kernel Name" in r][0]
k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
t = collections.Counter(); n = collections.Counter()
for r in rows:
    if len(r) == len(hdr) and r[hdr.index("Metric Name")] == "gpu__time_duration.sum":
        name = r[k].split("(")[0][:60]; t[name] += float(r[v].replace(",", "")); n[name] += 1
tot = sum(t.values())
for name, ns in t.most_common(8): print(f"{ns/1e6:10.3f} ms {n[name]:4d}x {100*ns/tot:5.1f}%  {name}")
PY
