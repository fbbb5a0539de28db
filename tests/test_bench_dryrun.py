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
    monkeypatch.setattr(bench, "_host_cores", lambda: 2)
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

    def fake_run(self, descs, scalars, domain, *, stream=None, subbox=None, halo_wait=None):
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
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" or not __import__("conftest").HAVE_GT4PY
    if extra != ["--no-autotune", "--no-pipeline"]:
        assert len(worker_calls) == (0 if extra else 1), line["config"]["autotune_top5"]
        assert isinstance(line["config"]["autotune_top5"], list) and line["config"]["autotune_candidates"] >= 10
        assert line["config"]["autotune_rejected"] == []
        assert "host pipeline" in line["e2e"]["path"] or "note" in line["e2e"]
        # the K-slab pipeline launched the stencil on sub-domains
        assert any(d[2] < 6 for d, _ in launched)
    else:
        assert line["config"]["autotune_top5"] is None and line["e2e"]["path"].startswith("serial")


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


TMA_WINNER = ["--options", '{"interior_loop": true, "static_pitch": 96, "tma": 2, "tile_j": 16, "prefetch": 1, "tma_mode": "bulk"}']


@pytest.mark.parametrize("extra", [[], ["--step-mode", "thin"], ["--no-overlap"], TMA_WINNER])
def test_bench_main_dry_run_two_ranks(monkeypatch, capsys, extra):
    """the N>1 control flow of bench.py (schedule self-check, trial, matched step counts) with the device,
    torch.distributed and the halo exchanger stubbed: rank 0 of a 2-rank job"""
    import torch
    import torch.distributed as dist

    import bench
    from gt4py_b200 import distributed, runtime, storage

    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(bench, "NI", 64)
    monkeypatch.setattr(bench, "NJ", 160)
    monkeypatch.setattr(bench, "NK", 4)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: _FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(storage, "_device", lambda device=None: torch.device("cpu"))
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda *a, **kw: real_tensor(*a, **{k: v for k, v in kw.items() if k != "device"}))
    monkeypatch.setattr(dist, "init_process_group", lambda *a, **kw: None)
    monkeypatch.setattr(dist, "barrier", lambda *a, **kw: None)
    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **kw: None)
    calls = {"exchange": 0, "subboxes": [], "streams": set()}

    class FakeExchanger:
        stream = 0x77

        def __init__(self, decomp, local_nj):
            assert decomp.n_ranks == 2 and local_nj == 160

        def exchange(self, fields):
            calls["exchange"] += 1
            return 3

        def close(self):
            pass

    monkeypatch.setattr(distributed, "HaloExchanger", FakeExchanger)

    class FakePeerHalo:  # symmetric memory needs real devices: same surface, plain storages
        stream = 0x88

        def __init__(self, decomp, local_nj):
            assert decomp.n_ranks == 2 and local_nj == 160
            self.epoch = 0

        def from_array(self, data, *, aligned_index=None, dimensions=None):
            return storage.from_array(data, aligned_index=aligned_index, dimensions=dimensions)

        def push(self, fields):
            self.epoch += 1
            calls["push"] = calls.get("push", 0) + 1
            return 2

        def wait_args(self):
            return (0x10, 0, self.epoch)

        def close(self):
            pass

    monkeypatch.setattr(distributed, "PeerHalo", FakePeerHalo)

    class FakeLib:
        def __getattr__(self, name):
            def fn(*a):
                import ctypes

                if name in ("b200_event_create", "b200_stream_create_priority"):
                    ctypes.cast(a[0], ctypes.POINTER(ctypes.c_void_p))[0] = 0x1000 + len(calls["streams"]) + 1
                    calls["streams"].add(name + str(len(calls["streams"])))
                return 0

            return fn

    monkeypatch.setattr(runtime, "load_library", lambda *a, **k: FakeLib())

    def fake_run(self, descs, scalars, domain, *, stream=None, subbox=None, halo_wait=None):
        calls["subboxes"].append((tuple(subbox) if subbox is not None else None, self.options.get("tile_j"), stream))
        if halo_wait is not None:
            assert self.options.get("halo_wait") and subbox is None and halo_wait[2] >= 1
            calls["halo_wait"] = calls.get("halo_wait", 0) + 1
        return len(self.plan["kernels"])

    monkeypatch.setattr(runtime.CompiledStencil, "run_descs", fake_run)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3", "--no-autotune", "--no-pipeline", *extra])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and "cpu_baseline" not in line and calls["exchange"] > 6
    par = line["config"]["parallelism"]
    thin_strips = [c for c in calls["subboxes"] if c[0] in ((0, 64, 0, 16), (0, 64, 144, 160))]
    if extra == TMA_WINNER:
        # a bulk-async winner: the peer schedule exists twice — with the winner's own halo_wait form (cross-proxy fence) and
        # with a register-window kernel; both were built at the register cap bench.py timed fastest
        trial = line["config"]["schedule_trial"]["ms_per_step"]
        assert set(trial) >= {"serial", "overlap", "thin", "peer", "peer_tma"}
        picks = line["config"]["halo_wait_kernels"]
        assert set(picks) == {"peer", "peer_tma"} and all("bare_kernel_ms" in p for p in picks.values())
        return
    if extra == ["--no-overlap"]:
        assert "then the whole slab" in par and not thin_strips and line["gpu_launches"] == 3 * (3 + 1)
    elif extra:
        assert "16-row boundary strips" in par and thin_strips and line["gpu_launches"] == 3 * (3 + 3)
        assert all(tj == 16 and stream is not None for _, tj, stream in thin_strips)  # short-tile kernel, side stream
        assert ((0, 64, 16, 144), None, None) in calls["subboxes"]  # interior: the tuned kernel on the compute stream
    else:
        trial = line["config"]["schedule_trial"]["ms_per_step"]
        assert set(trial) >= {"serial", "overlap", "thin", "peer"} and calls["push"] == calls["halo_wait"] > 3
    assert "exposed_comm_us_per_step" in line["config"] and line["config"]["multi_gpu_check"] is not None


def test_watchdog_prints_the_device_timed_part(monkeypatch, capsys):
    """a hang in a later, optional phase (end-to-end pipeline) must not lose the device-timed measurement"""
    import os

    import bench

    monkeypatch.setattr(os, "_exit", lambda code: (_ for _ in ()).throw(SystemExit(code)))
    monkeypatch.setattr(bench, "_PARTIAL_LINE", {"metric": bench.METRIC, "value": 1.0, "e2e": {"value": None, "note": "not measured"}})
    with pytest.raises(SystemExit) as exc:
        bench._watchdog_abort(5)
    assert exc.value.code == 3
    out = capsys.readouterr()
    assert json.loads(out.out.strip())["value"] == 1.0 and "watchdog fired" in out.err
    monkeypatch.setattr(bench, "_PARTIAL_LINE", None)
    with pytest.raises(SystemExit):
        bench._watchdog_abort(5)
    assert capsys.readouterr().out == ""
