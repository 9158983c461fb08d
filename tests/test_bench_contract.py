"""The JSON lines bench.py prints (native arm and reference arm) carry every key of the driver's contract.
Checked on the committed lines of the last GPU round (profiles/), so it runs without a GPU."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    with open(files[-1]) as f:
        return json.loads([l for l in f.read().splitlines() if l.startswith("{")][-1])


def test_native_line_has_the_contract_keys():
    d = _latest("r2_v*_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] > 0 and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # every BASELINE config rides in the one line (round 2)
    h = d["hour_sharded"]
    assert h["segments"] == 720 and h["scaling"] == "strong" and h["ms"] > 0 and abs(h["value"] - 3600.0 / (h["ms"] / 1e3)) < 1e-3 * h["value"]
    f = d["fusion_only_65536"]
    assert f["rows"] == 65536 and f["ms"] > 0 and 0 < f["frac"] < 1 and abs(f["issued_bf16_tflops"] - 3 * f["algorithmic_tflops"]) < 1e-6
    assert d["stream_latency"]["p50_ms"] > 0 and d["reference_api_call"]["p50_ms"] > 0
    assert 0 < r["fp32"]["frac"] < 1 and r["msa_version"] >= 200
    assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes_per_launch"]


def test_reference_line_has_the_contract_keys():
    d = _latest("r2_v*_bench_reference.json")
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "audio-s/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    n = _latest("r2_v*_bench.json")
    assert d["metric"] == n["metric"] and d["unit"] == n["unit"] and d["higher_is_better"] == n["higher_is_better"]
