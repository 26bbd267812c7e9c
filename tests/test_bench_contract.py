"""The bench.py contract: the reference arm really runs here (CPU), and the committed GPU lines carry every key
the driver reads (profiles/r01_bench_*.json are the lines bench.py printed on the B200 box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def last_json_line(text):
    lines = [l for l in text.splitlines() if l.startswith("{")]
    assert lines, text[-2000:]
    return json.loads(lines[-1])


def test_reference_arm_runs_on_the_host_cores():
    from oracle.oracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libspmvref.so not built")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = last_json_line(p.stdout)
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"] == "spmv_effective_bandwidth" and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["config"]["workload"] == "c5_csr" and line["dtype"] == "f64" and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] > 0
    # config 5 does not fit the reference's int32 sizes: the CPU sample is a 512 x 512 x 24 slab of the same operator,
    # and the line says so: 12 * 1534^2 * 70 + 4 * (rows + 1) + 16 * rows
    assert cb["algorithmic_bytes"] == 12 * 1534 * 1534 * 70 + 4 * (512 * 512 * 24 + 1) + 16 * 512 * 512 * 24
    assert "SLAB" in cb["sample"] and "slab" in line["config"]["sample"].lower()
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.parametrize("name", ["r01_bench_n1.json", "r01_bench_n2.json", "r01_bench_n8.json", "r01_bench_c4_hyb_n2.json"])
def test_committed_gpu_lines_carry_the_contract(name):
    line = last_json_line(open(os.path.join(ROOT, "profiles", name)).read())
    assert BASE_KEYS <= set(line), sorted(BASE_KEYS - set(line))
    assert line["metric"] == "spmv_effective_bandwidth" and line["unit"] == "GB/s" and line["dtype"] == "f64"
    assert line["data"] == "synthetic" and line["vs_baseline"] is None and line["higher_is_better"] is True
    assert "workload" in line["config"] and "model" not in line["config"]
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = line["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0
    assert e["value"] < line["value"]  # host buffers cross PCIe every step: never the device-resident number
    assert line["gpu_launches"] > 0
    c = line["clocks"]
    assert c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if line["n_gpus"] == 1:
        assert line["scaling"] == "weak" and line["config"]["workload"] == "c2_ell"
        cb = line["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("reference", "port")
        assert r["traffic"] and 0.95 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.1
        assert {f["workload"] for f in line["formats"]} >= {"c1_csr", "c1_ell", "c1_coo", "c1_hyb", "c3_coo", "c4_hyb", "c5_csr"}
    else:
        assert line["scaling"] == "strong" and line["single_gpu"]["speedup"] > 1.0


@pytest.mark.parametrize("name", ["r02_bench_n1.json", "r02_bench_n2.json", "r02_bench_n4.json", "r02_bench_n8.json"])
def test_round2_gpu_lines_are_one_workload_with_parity(name):
    """Round 2: every GPU count runs BASELINE config 5 (so the driver's 1 -> 8 curve is one workload), every line was
    preceded by a parity check in the same run, and the multi-GPU lines come from the executor below the C ABI."""
    line = last_json_line(open(os.path.join(ROOT, "profiles", name)).read())
    assert BASE_KEYS <= set(line), sorted(BASE_KEYS - set(line))
    assert line["metric"] == "spmv_effective_bandwidth" and line["unit"] == "GB/s" and line["dtype"] == "f64"
    assert line["config"]["workload"] == "c5_csr" and line["config"]["algorithmic_bytes"] == 46001250212
    assert line["scaling"] == "strong" and line["vs_baseline"] is None and "model" not in line["config"]
    p = line["parity"]
    assert p["ok"] is True and p["bad_rows"] == 0 and p["rows_checked"] >= 1_000_000
    r = line["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.2
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < line["value"]
    assert line["gpu_launches"] > 0
    c = line["clocks"]
    assert c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    n = line["n_gpus"]
    if n == 1:
        cb = line["cpu_baseline"]
        assert cb["kind"] in ("reference", "port") and "SLAB" in cb["sample"]
        t = line["targets"]
        for k in ("c1_csr", "c1_ell"):
            assert t[k]["frac_of_8TBs_pipelined"] >= 0.70  # BASELINE's target, sustained
            assert t[k]["xy_bytes_in_cycle"] >= 2 * 126e6  # x + y of the rotating copies exceed twice the L2
            assert t[k]["us_isolated"] > t[k]["us_pipelined"]
        assert {f["workload"] for f in line["formats"]} >= {"c1_coo", "c1_hyb", "c2_ell", "c2_csr", "c3_coo", "c4_hyb"}
    else:
        assert "spmvb200_dist" in line["config"]["executor"]
        assert line["single_gpu"]["speedup"] > 0.9 * n  # halo plan: 1.98x / 3.95x / 7.83x
        v = line["exchange_variants"]
        assert v["halo"]["parity"]["ok"] and v["allgather"]["parity"]["ok"] and v["halo"]["parity"]["ranks"] == n
        assert v["halo"]["recv_bytes_per_step_per_rank"] in (2 * 512 * 512 * 8, 512 * 512 * 8)
        if n == 2:
            h = line["c4_hyb"]
            assert h["parity"]["ok"] and h["single_gpu"]["speedup"] > 1.5 and h["nonzeros"] == 2103842462


@pytest.mark.parametrize("name", ["r03_bench_n1.json", "r03_bench_n2.json", "r03_bench_n4.json", "r03_bench_n8.json"])
def test_round2_final_lines_with_the_diagonal_slices(name):
    """The final lines of round 2: the sliced CSR kernel stores its column stream by diagonal, so the effective bandwidth
    (ALGORITHMIC bytes / time) exceeds the measured peak -- the line must then carry the measured traffic, a physical DRAM
    rate that does not, and the time of the same steps with every index stored."""
    line = last_json_line(open(os.path.join(ROOT, "profiles", name)).read())
    assert BASE_KEYS <= set(line), sorted(BASE_KEYS - set(line))
    assert line["config"]["workload"] == "c5_csr" and line["config"]["algorithmic_bytes"] == 46001250212
    assert line["parity"]["ok"] is True and line["parity"]["bad_rows"] == 0 and line["parity"]["rows_checked"] >= 1_000_000
    r = line["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < line["value"]
    c = line["clocks"]
    assert c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    n = line["n_gpus"]
    if n == 1:
        ir = r["index_runs"]
        assert ir["active"] is True and ir["column_indices_stored"] < 0.1 * ir["stored_entries"]
        assert r["traffic"] < r["algorithmic_bytes_per_launch"]             # fewer bytes moved than the metric counts ...
        assert r["dram_gbs"] < 1.05 * r["peak"] < r["achieved"]              # ... at a physical rate the memory can deliver
        assert abs(r["dram_gbs"] - r["traffic"] / (r["kernel_ms_mean"] * 1e-3) / 1e9) < 1e-6
        assert ir["ms_per_step_with_every_index_stored"] > 1.3 * line["ms_per_step"]
        assert ir["traffic_with_every_index_stored"] > r["algorithmic_bytes_per_launch"]
        assert line["config"]["resident_bytes"] < 0.75 * 46001250212
        for k in ("c1_csr", "c1_ell"):
            assert line["targets"][k]["frac_of_8TBs_pipelined"] >= 0.70
    else:
        assert line["single_gpu"]["speedup"] > 0.9 * n
        v = line["exchange_variants"]
        assert all(v[k]["parity"]["ok"] for k in ("allgather", "allgather_peer", "halo", "halo_peer", "halo_push"))
        if n == 2:
            assert line["c4_hyb"]["parity"]["ok"] and line["c4_hyb"]["single_gpu"]["speedup"] > 1.5
