"""CPU-side checks of the measurement tooling: workload plans of the non-headline configs
(tools/bench_workloads.py) and the reference arm of bench.py."""

import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))


def test_workload_plans_match_the_baseline_configs():
    import bench_workloads as bw

    p = bw.plan("upwind5", 8)
    assert p["domain_per_gpu"] == (2048, 256, 80) and p["scaling"] == "strong" and p["bytes_per_cell"] == [16]
    assert p["exchanges"] == {"0": [("phi", 3)]}
    p = bw.plan("fastwaves", 8)
    assert p["domain_per_gpu"] == (4096, 512, 80) and p["scaling"] == "weak" and sum(p["bytes_per_cell"]) == 72
    # the divergence reads what the pressure-gradient step wrote; the implicit solve reads the divergence
    assert {"u", "v", "div", "pp", "pp_new", "w"} <= set(p["buffers"])
    p = bw.plan("tridiagonal", 1)
    assert p["domain_per_gpu"] == (512, 512, 160) and p["bytes_per_cell"] == [56] and p["exchanges"] == {}


def test_workload_cpu_baseline_runs_the_reference_numpy_backend_or_the_oracle_chain(monkeypatch):
    import bench_workloads as bw

    w = bw.workload("fastwaves", 1)
    r = bw.cpu_baseline(bw.step_description(w), w["halo"], sample=(12, 10, 6))
    # the reference's own numpy backend where the reference package is importable (baseline/_ref or /root/reference) ...
    assert r["value"] > 0 and r["kind"] in ("reference", "port") and r["cores"] == 1
    if r["kind"] == "reference":
        assert "reference gt4py numpy backend" in r["sample"]
    # ... else (simulated here) the oracle port
    monkeypatch.setenv("B200_NO_REFERENCE", "1")
    r = bw.cpu_baseline(bw.step_description(w), w["halo"], sample=(12, 10, 6))
    assert r["value"] > 0 and r["kind"] == "port" and r["sample"].startswith("oracle")


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)  # fmt: skip
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    # the reference's own numpy backend where the reference package is importable (baseline/_ref or /root/reference), else the port
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert "1024x1024x80" in line["config"]["workload"] and "whole 1024x1024x80" in line["cpu_baseline"]["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "Mcell-updates/s"
