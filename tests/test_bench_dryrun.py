"""bench.py's main path executed on the CPU with the device stubbed out (launches, CUDA events,
streams, pinned memory): catches control-flow / bookkeeping errors in the script the driver runs,
which otherwise only a GPU box would reveal.  No timing or parity claim is made here."""

import contextlib
import json
import sys

import numpy as np
import pytest


class _FakeEvent:
    clock = 0.0

    def __init__(self, enable_timing=False):
        self.at = 0.0

    def record(self, stream=None):
        _FakeEvent.clock += 0.25
        self.at = _FakeEvent.clock

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return max(other.at - self.at, 0.01)


class _FakeStream:
    cuda_stream = 0

    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass


@pytest.mark.parametrize("extra", [[], ["--tune-in-process"], ["--no-autotune", "--no-pipeline"]])
def test_bench_main_dry_run(monkeypatch, capsys, extra):
    import torch

    import bench
    from gt4py_b200 import runtime, storage

    monkeypatch.setattr(bench, "NI", 64)
    monkeypatch.setattr(bench, "NJ", 48)
    monkeypatch.setattr(bench, "NK", 6)
    monkeypatch.setattr(bench, "CPU_SAMPLE", (16, 16, 4))
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: _FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", _FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(storage, "_device", lambda device=None: torch.device("cpu"))
    launched = []

    def fake_run(self, descs, scalars, domain, *, stream=None, subbox=None):
        launched.append((tuple(domain), dict(self.options)))
        return len(self.plan["kernels"])

    monkeypatch.setattr(runtime.CompiledStencil, "run_descs", fake_run)

    # the autotune child (`python -m gt4py_b200.tune_worker spec.json`) runs in-process under the same stubs
    import io
    import subprocess

    real_run = subprocess.run
    worker_calls = []

    def fake_subprocess_run(cmd, *a, **kw):
        if "gt4py_b200.tune_worker" not in cmd:
            return real_run(cmd, *a, **kw)
        from gt4py_b200 import tune_worker

        assert kw.get("timeout") and "RANK" not in kw["env"]
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            rc = tune_worker.main([cmd[-1]])
        worker_calls.append(cmd)
        return subprocess.CompletedProcess(cmd, rc, stdout=buf.getvalue(), stderr="")

    monkeypatch.setattr(subprocess, "run", fake_subprocess_run)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "3", "--warmup", "3", "--watchdog", "300", *extra])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "gpu_launches", "e2e", "roofline", "clocks", "cpu_baseline"):  # fmt: skip
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["gpu_launches"] == 3 and line["vs_baseline"] is None
    assert line["roofline"]["bound"] == "hbm" and 0 < line["roofline"]["frac"] and line["roofline"]["peak"] > 1000
    assert line["e2e"]["h2d_bytes_per_step"] == 2 * line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    if extra != ["--no-autotune", "--no-pipeline"]:
        assert len(worker_calls) == (0 if extra else 1), line["config"]["autotune"]
        assert isinstance(line["config"]["autotune"], list) and len(line["config"]["autotune"]) >= 10
        assert line["config"]["autotune_rejected"] == []
        assert "host pipeline" in line["e2e"]["path"] or "note" in line["e2e"]
        # the K-slab pipeline launched the stencil on sub-domains
        assert any(d[2] < 6 for d, _ in launched)
    else:
        assert line["config"]["autotune"] is None and line["e2e"]["path"].startswith("serial")


@pytest.mark.parametrize("workload", ["tridiagonal", "upwind5", "fastwaves", "hdiff_x2", "hdiff_x2 --fuse"])
def test_workload_bench_dry_run(monkeypatch, capsys, workload):
    import pathlib

    import torch

    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent / "tools"))
    import bench_workloads as bw
    from gt4py_b200 import runtime, storage

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: _FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(storage, "_device", lambda device=None: torch.device("cpu"))
    monkeypatch.setattr(runtime.CompiledStencil, "run_descs", lambda self, d, s, dom, **kw: len(self.plan["kernels"]))
    monkeypatch.setattr(bw, "cpu_baseline", lambda steps, halo, **kw: {"value": 1.0, "unit": "Mcell-updates/s", "cores": 1, "kind": "port", "sample": "stub"})
    monkeypatch.setattr(sys, "argv", ["bench_workloads.py", "--workload", *workload.split(), "--steps", "3", "--shrink", "64"])
    bw.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["n_gpus"] == 1 and line["gpu_launches"] > 0 and line["roofline"]["bytes_per_cell"] == {"tridiagonal": 56, "upwind5": 16, "fastwaves": 72, "hdiff_x2": 24, "hdiff_x2 --fuse": 12}[workload]
    if workload.endswith("--fuse"):
        assert line["gpu_launches"] == 3 and len(line["config"]["kernels"]) == 1  # one launch per pass of two updates
    assert "SMOKE RUN" in line["config"]["workload"] and line["cpu_baseline"]["kind"] == "port"
