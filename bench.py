"""bench.py — headline benchmark of the b200 stencil backend (contract: see the task statement).

Workload (BASELINE.json configs[1]): horizontal diffusion (lap-of-lap + flux limiter,
reference stencil tests/cartesian_tests/integration_tests/multi_feature_tests/stencil_definitions.py:316-328)
on a 1024 x 1024 x 80 fp32 domain (fields 1028 x 1028 x 80, origin (2,2,0)), compiled with
literal_float_precision=32.  One "step" = one application of the stencil over the whole domain.

  value  : Mcell-updates/s with all fields resident in HBM, timed with CUDA events on the
           launching stream over exactly K steps (max over ranks for N > 1)
  e2e    : same metric through the public call (`stencil(in, out, coeff, origin=…, domain=…)`)
           with HOST (pinned) buffers: H2D copy of in_field+coeff and D2H copy of out_field
           inside the timed region every step
  roofline: algorithmic bytes (12 B/cell: read in_field, coeff; write out_field) over
           the kernel's average launch duration, against the measured HBM copy bandwidth of
           MEASURED_PEAKS.json
  N > 1  : weak scaling — the global domain is 1024 x (1024*N) x 80, cut into N J-slabs; every step
           exchanges the two 2-row J-halos of in_field with the neighbour ranks over NCCL (C-ABI
           b200_halo_exchange).  Three step schedules (exchange then whole slab / exchange overlapped
           with the interior + whole-tile boundary strips / + thin strips on a high-priority stream)
           are checked bit for bit against each other on every rank, timed briefly, and the fastest
           on the max over ranks is used (`--step-mode` forces one).

The code-generation variant is autotuned in a sacrificial child process (gt4py_b200/tune_worker.py);
every variant must reproduce the default one bit for bit before it is timed.

`--impl reference` times the CPU restatement of the reference numpy backend (oracle/) on a bounded
sample of the same workload, one process per usable host core (the reference's GridTools CPU backends
cannot be built offline: gridtools-cpp headers are not vendored).
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

NI, NJ, NK, HALO = 1024, 1024, 80, 2
STENCIL, VARIANT = "hdiff_f32", "staged"
BYTES_PER_CELL = 12
METRIC = "Mcell-updates/s + achieved HBM GB/s, horiz-diffusion 1024x1024x80 fp32"
WORKLOAD = (f"horizontal diffusion (lap-of-lap + flux limiter) {NI}x{NJ}x{NK} fp32 per GPU (BASELINE configs[1]), "
            f"literal_float_precision=32, fields {(NI + 2 * HALO, NJ + 2 * HALO, NK)} origin {(HALO, HALO, 0)}")
_PARTIAL_LINE = None  # rank 0: the bench line without e2e, once the device-timed part is done (printed by the watchdog)
STRIP = 64  # rows of the boundary strips in the overlapped multi-GPU step (reset to the tuned J tile of the kernel)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- CPU arm: the reference's own `numpy` backend --------------------------------------------------------------
# gt4py (the unmodified reference package, baseline/_ref, tools/install_reference.sh) builds the SAME GTScript
# definition (tools/stencil_defs.py: hdiff_f32) with backend="numpy" and runs it through StencilObject.__call__.
# The numpy backend executes whole-array NumPy statements: one thread.  To give the host "all the threads it can use"
# the 1024 x 1024 x 80 domain is cut into K slabs (horizontal diffusion has no vertical coupling), one slab per usable
# core, every process running the reference stencil on its slab, started together: one step = the whole domain once.
# Falls back to the oracle port (oracle/numpy_oracle.py, kind "port") only where the reference is not importable.
def _host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _enable_reference() -> bool:
    sys.path.insert(0, str(ROOT / "tools"))
    import refenv

    return refenv.enable_gt4py()


def _reference_stencil():
    """hdiff_f32 compiled by the reference's numpy backend (cached in GT_CACHE_ROOT by gt4py itself)."""
    import warnings

    warnings.filterwarnings("ignore")
    import stencil_defs
    from gt4py.cartesian import gtscript

    case = stencil_defs.REGISTRY[STENCIL]
    return gtscript.stencil(backend="numpy", definition=case["definition"], externals=case["externals"] or {},
                            name=f"{STENCIL}_bench_ref", **case["build"])  # fmt: skip


def _slab_levels(n_procs: int):
    base, rem = divmod(NK, n_procs)
    return [base + (1 if r < rem else 0) for r in range(n_procs)]


def _cpu_worker(kind, barrier, steps, warmup, seed, out_q, dims):
    """one process of the CPU arm: `steps` calls on a (NI, NJ, nk) slab of the workload"""
    NI, NJ, nk = dims  # (passed explicitly: the workers are fresh interpreters)
    rng = np.random.default_rng(seed)
    shape = (NI + 2 * HALO, NJ + 2 * HALO, nk)
    origin = (HALO, HALO, 0)
    fields = {"in_field": rng.random(shape, dtype=np.float32), "coeff": rng.random(shape, dtype=np.float32) * np.float32(0.1),
              "out_field": np.zeros(shape, np.float32)}  # fmt: skip
    if kind == "reference":
        _enable_reference()
        st = _reference_stencil()
        call = lambda: st(**fields, origin=origin, domain=(NI, NJ, nk))  # noqa: E731
    else:
        from gt4py_b200 import testing
        from oracle import numpy_oracle

        ir = testing.load_ir(STENCIL, "default")
        origins = {n: origin for n in fields}
        call = lambda: numpy_oracle.run(ir, fields, {}, (NI, NJ, nk), origins)  # noqa: E731
    for _ in range(max(1, warmup)):
        call()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    out_q.put(time.perf_counter() - t0)


def cpu_throughput(kind: str, n_procs: int, steps: int, warmup: int = 1, levels=None):
    """-> (Mcell-updates/s over n_procs processes that together cover `levels` K levels per step, seconds of the slowest)"""
    levels = levels if levels is not None else _slab_levels(n_procs)
    if n_procs <= 1:
        import queue

        class _NoBarrier:
            def wait(self):
                pass

        q = queue.Queue()
        _cpu_worker(kind, _NoBarrier(), steps, warmup, 0, q, (NI, NJ, levels[0]))
        dt = q.get()
        return NI * NJ * levels[0] * steps / dt / 1e6, dt
    import multiprocessing as mp

    # forkserver: the workers are forked from a clean single-threaded server process, never from this one
    # (which runs a watchdog thread and, later, owns a CUDA context)
    ctx = mp.get_context("forkserver")
    barrier, q = ctx.Barrier(n_procs), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(kind, barrier, steps, warmup, r, q, (NI, NJ, levels[r])), daemon=True) for r in range(n_procs)]
    for pr in procs:
        pr.start()
    try:
        times = [q.get(timeout=240) for _ in procs]
        for pr in procs:
            pr.join(timeout=30)
    finally:
        for pr in procs:  # a worker that died or hangs must not outlive the measurement
            if pr.is_alive():
                pr.terminate()
    dt = max(times)
    return NI * NJ * sum(levels) * steps / dt / 1e6, dt


def cpu_arm(steps: int, warmup: int, max_seconds: float):
    """The CPU arm on the full 1024 x 1024 x 80 workload: -> (cpu_baseline dict, ms per step of the whole domain)."""
    kind = "reference"
    try:
        if not _enable_reference():
            raise ImportError("gt4py not importable")
        _reference_stencil()  # build once here: the workers find it in gt4py's cache
    except Exception as exc:  # the reference is not here: time the port, say so
        kind = f"port ({type(exc).__name__}: {str(exc)[:80]})"
    k = "reference" if kind == "reference" else "port"
    cores = min(_host_cores(), NK)
    single, dt1 = cpu_throughput(k, 1, 1, 1, levels=[max(1, NK // cores)])
    # bound the run: a step of the whole domain takes about (cells / multi-process throughput); assume no better than linear
    est_step = NI * NJ * NK / (single * cores * 1e6)
    steps = int(max(1, min(steps, (max_seconds - 2 * dt1) / max(est_step * 1.6, 1e-3) - warmup)))
    try:
        multi, dt = cpu_throughput(k, cores, steps, max(1, min(warmup, 2))) if cores > 1 else (single, dt1)
    except Exception:
        multi, dt, cores = single, dt1, 1
    if multi < single:
        multi, cores = single, 1
    ms_step = NI * NJ * NK / (multi * 1e6) * 1e3
    line = {
        "value": round(multi, 3), "unit": "Mcell-updates/s", "cores": cores, "kind": k, "single_core_value": round(single, 3),
        "sample": (f"{'reference gt4py numpy backend (baseline/_ref), StencilObject.__call__' if k == 'reference' else 'oracle port (' + kind + ')'}"
                   f" on the whole {NI}x{NJ}x{NK} fp32 domain per step, cut into {cores} K slabs of {NK // cores}-{-(-NK // cores)} levels, one "
                   f"process per usable core started together ({os.cpu_count()} host cores present), {steps} timed steps"),
        # north_star: "... and gt:cpu_kfirst where it builds offline" — it does not: its generated C++ needs the GridTools
        # headers of the gridtools_cpp package, which is not in this image
        "gt_cpu_kfirst": "not available: the reference's gt:cpu_kfirst backend needs the gridtools_cpp C++ headers, absent offline",
    }  # fmt: skip
    return line, ms_step, steps


def cpu_baseline(max_seconds: float = 25.0):
    """Bounded sample for the default bench line."""
    return cpu_arm(3, 1, max_seconds)[0]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms_step, steps = cpu_arm(max(1, args.steps), max(1, min(args.warmup, 2)), max_seconds=150.0)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": base["value"],
        "unit": "Mcell-updates/s",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": args.warmup,
        "ms_per_step": round(ms_step, 3),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---- clocks sampling ---------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thr = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip().splitlines()
                if out:
                    self.samples.append([x.strip() for x in out[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons,
            "samples": len(self.samples),
        }


def _watchdog_abort(seconds: float):
    """The run exceeded its time limit: print what was measured so far (rank 0, if the device-timed part is done) and leave."""
    if _PARTIAL_LINE is not None:
        print(json.dumps(_PARTIAL_LINE), flush=True)
    sys.stderr.write(f"bench.py: watchdog fired after {seconds:.0f} s (rank {os.environ.get('RANK', '0')}) - aborting\n")
    sys.stderr.flush()
    os._exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--strategy", default="auto")
    ap.add_argument("--overlap", action="store_true", help="N>1: always overlap the halo exchange with the interior tiles")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: always exchange first, then compute the whole slab")
    ap.add_argument("--no-peer", action="store_true", help="N>1: do not use the peer-memory (symmetric memory) halo exchange")
    ap.add_argument("--step-mode", default=None, choices=["serial", "overlap", "thin", "peer", "peer_tma"],
                    help="N>1: force one step schedule (default: a short trial of all of them, fastest on the max over ranks wins)")
    ap.add_argument("--watchdog", type=float, default=600.0, help="abort instead of hanging after this many seconds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-autotune", action="store_true", help="use the default code-generation options")
    ap.add_argument("--options", default=None, help="JSON dict of code-generation options to use as they are (implies --no-autotune): "
                    "how the ncu capture of the variant a bench run selected is taken")
    ap.add_argument("--tune-in-process", action="store_true", help="run the autotune sweep in this process instead of a child")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e: serial copy-in / call / copy-out only")
    ap.add_argument("--pipeline-chunks", type=int, default=10, help="K slabs of the host pipeline (e2e)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # N>1 default: time a few steps of both orders after the warm-up and keep the faster one (decided on
    # the max over ranks, so every rank takes the same decision)
    args.step_mode = args.step_mode or ("overlap" if args.overlap else ("serial" if args.no_overlap else "auto"))

    # never hang the box: a lost peer / unmatched exchange turns into a loud non-zero exit
    wd = threading.Timer(args.watchdog, _watchdog_abort, args=(args.watchdog,))
    wd.daemon = True
    wd.start()
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        wd.cancel()


def run_b200(args):
    import torch

    from gt4py_b200 import runtime, storage, testing
    from gt4py_b200.distributed import HaloExchanger, PeerHalo, SlabDecomposition
    from gt4py_b200.stencil import B200Stencil

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the b200 backend has no CPU fallback")
    # host baseline first: its worker processes are forked before this process owns a CUDA context
    cpu_line = cpu_baseline() if (world == 1 and not args.no_cpu_baseline) else None
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world

    # N > 1: exchanged fields live in symmetric memory so that the neighbours can store their boundary rows straight
    # into this rank's halo rows (PeerHalo: one push kernel per step, flags consumed inside the stencil kernel)
    peer, peer_note = None, None
    if world > 1 and not args.no_peer:
        try:
            peer = PeerHalo(SlabDecomposition(world, rank, NJ * world), NJ)
        except Exception as exc:
            peer_note = f"peer-memory exchange unavailable ({type(exc).__name__}: {str(exc)[:160]})"
        okp = torch.tensor([1.0 if peer is not None else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(okp, op=dist.ReduceOp.MIN)  # symmetric allocations are collective: all ranks or none
        if float(okp.item()) < 0.5:
            peer = None
    st_ir = testing.load_ir(STENCIL, VARIANT)
    forced = json.loads(args.options) if args.options else {}
    if forced:
        args.no_autotune = True
    stencil = B200Stencil(st_ir, {"strategy": args.strategy, "device_sync": False, **forced})
    shape = (NI + 2 * HALO, NJ + 2 * HALO, NK)
    origin3 = (HALO, HALO, 0)
    origins = {"in_field": origin3, "out_field": origin3, "coeff": origin3}
    domain = (NI, NJ, NK)
    rng = np.random.default_rng(rank)
    # two buffer sets, rotated every step (working set 2 x 1 GB >> 126 MB L2)
    sets = []
    host_in = host_coeff = None
    for s in range(2):
        h_in = rng.random(shape, dtype=np.float32)
        h_co = rng.random(shape, dtype=np.float32) * np.float32(0.1)
        if peer is not None:
            sets.append({"in_field": peer.from_array(h_in, aligned_index=origin3), "coeff": storage.from_array(h_co, aligned_index=origin3),
                         "out_field": storage.zeros(shape, np.float32, aligned_index=origin3)})  # fmt: skip
        else:
            sets.append(
                {
                    "in_field": storage.from_array(h_in, aligned_index=origin3),
                    "coeff": storage.from_array(h_co, aligned_index=origin3),
                    "out_field": storage.zeros(shape, np.float32, aligned_index=origin3),
                }
            )
        if s == 0:
            host_in, host_coeff = h_in, h_co
    tuned = None
    if not args.no_autotune:
        # pick the fastest code-generation variant for this stencil x domain on this device (every
        # candidate is checked bit-for-bit against the default variant before it is timed)
        # The sweep runs in a sacrificial child process (gt4py_b200/tune_worker.py): a variant that faults
        # or hangs on the device cannot take this process' CUDA context, and with it the bench line.
        try:
            if args.tune_in_process:
                tuned = stencil.autotune(sets[0], {}, domain=domain, origin=origins, iters=20)
            else:
                tuned = stencil.autotune_isolated(sets[0], {}, domain=domain, origin=origins, iters=20, timeout=min(180.0, args.watchdog / 3), device=local_rank)
        except Exception as exc:  # keep the measured default rather than lose the bench line
            tuned = f"autotune failed, default options used: {type(exc).__name__}: {str(exc)[-400:]}"
    frozen = stencil.freeze(origin=origins, domain=domain)
    global STRIP
    STRIP = int(stencil.backend_options.get("tile_j", 64))  # boundary strips = whole J tiles of the tuned kernel
    if 2 * STRIP >= NJ:
        STRIP = 64

    exchanger = None
    if world > 1:
        decomp = SlabDecomposition(world, rank, NJ * world)
        exchanger = HaloExchanger(decomp, NJ)
    lib = runtime.load_library()
    main_stream = torch.cuda.current_stream().cuda_stream
    ev_pool = []

    def make_event():
        import ctypes

        e = ctypes.c_void_p()
        runtime.check(lib.b200_event_create(ctypes.byref(e)))
        ev_pool.append(e)
        return e

    ev_ready, ev_halo, ev_strips = (make_event(), make_event(), make_event()) if world > 1 else (None, None, None)
    # "thin" schedule: the boundary rows are computed by a short-tile variant of the same kernel (THIN rows per
    # tile: a strip is one short wave instead of one full 64-row march) on a high-priority side stream, so they
    # run as soon as the halo has arrived, concurrently with the tail of the interior kernel
    THIN = 16
    thin_frozen, strip_stream = None, None
    if world > 1:
        import ctypes

        try:
            thin = B200Stencil(st_ir, {**stencil.backend_options, "tile_j": THIN, "device_sync": False})
            if thin.compiled.kernel_names() and all(k["kind"] == "stream" for k in thin.compiled.plan["kernels"]):
                thin_frozen = thin.freeze(origin=origins, domain=domain)
                h = ctypes.c_void_p()
                runtime.check(lib.b200_stream_create_priority(ctypes.byref(h), 1))
                strip_stream = int(h.value)
        except Exception:
            thin_frozen = None
        # the schedule list must be the same on every rank (the trial below runs matched exchanges)
        avail = torch.tensor([1.0 if thin_frozen is not None else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(avail, op=dist.ReduceOp.MIN)
        if float(avail.item()) < 0.5:
            thin_frozen = None

    hw_pick, hwt_pick = None, None

    def halo_wait_variant(base_opts):
        """The `halo_wait` form of a kernel variant, at the register cap that runs fastest: the flag-wait code costs 8-14
        registers (LDG 44 -> 54, TMA 50 -> 64: two resident CTAs per SM less, 0.169 -> 0.193 ms for the bare LDG kernel,
        r02u) which `__launch_bounds__` gives back without spills (0.182 ms at a 44-register cap, r02y).  Timed without
        flags (epoch 0: nobody waits)."""
        best = (None, None, float("inf"))
        threads = 32 * int(base_opts.get("warps", 4))
        caps = sorted({min(32, 2048 // threads, 65536 // (regs * threads)) for regs in (48, 44, 40)})  # resident CTAs per SM at that many registers
        for mb in (None, *caps):
            opts = {**base_opts, "halo_wait": True, "device_sync": False, **({"min_blocks": mb} if mb else {})}
            try:
                cand = B200Stencil(st_ir, opts)
                if not all(k["kind"] == "stream" for k in cand.compiled.plan["kernels"]):
                    continue
                fr = cand.freeze(origin=origins, domain=domain)
                for i in range(3):
                    fr(**sets[i & 1])
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(20):
                    fr(**sets[i & 1])
                b.record()
                b.synchronize()
                ms = a.elapsed_time(b) / 20
            except Exception:
                if mb is None:
                    raise
                continue
            if ms < best[2]:
                best = (fr, mb, ms)
        if best[0] is None:
            raise RuntimeError("no streaming halo_wait kernel")
        return best[0], {"min_blocks": best[1], "bare_kernel_ms": round(best[2], 5)}

    # "peer" schedule: ONE launch of the `halo_wait` variant of the tuned kernel per step (boundary tiles last, waiting on
    # the neighbours' flags on the device) + one push kernel on the comm stream; no NCCL call, no boundary strips
    peer_frozen, peer_tma_frozen = None, None
    if peer is not None:
        try:
            hw_opts = dict(stencil.backend_options)
            if hw_opts.get("tma"):
                # TMA variants are not combined with halo_wait: the fastest REGISTER-WINDOW candidate of the sweep instead
                # (not the TMA winner minus its ring: its short J tiles are tuned for the ring — 0.183 vs 0.168 ms, r02q)
                ldg = [c for c, _ms in (tuned if isinstance(tuned, list) else []) if not c.get("tma")]
                base = {k: v for k, v in stencil.backend_options.items() if k in ("strategy", "device_sync")}
                hw_opts = {**base, **(ldg[0] if ldg else {"interior_loop": True, "static_pitch": int(hw_opts.get("static_pitch", 0) or 0)})}
                if not hw_opts.get("static_pitch"):
                    hw_opts.pop("static_pitch", None)
            peer_frozen, hw_pick = halo_wait_variant(hw_opts)
        except Exception as exc:
            peer_note = f"halo_wait kernel unavailable ({type(exc).__name__}: {str(exc)[:160]})"
        # "peer_tma": the same schedule with the TMA winner itself as the waiting kernel (flag acquire in the generic proxy,
        # fence.proxy.async.global, then the bulk copies of the halo rows through the async proxy)
        if stencil.backend_options.get("tma"):
            try:
                peer_tma_frozen, hwt_pick = halo_wait_variant(stencil.backend_options)
            except Exception as exc:
                peer_note = f"halo_wait TMA kernel unavailable ({type(exc).__name__}: {str(exc)[:160]})"
        okp = torch.tensor([1.0 if peer_frozen is not None else 0.0, 1.0 if peer_tma_frozen is not None else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(okp, op=dist.ReduceOp.MIN)
        if float(okp[0].item()) < 0.5:
            peer_frozen = None
        if float(okp[1].item()) < 0.5:
            peer_tma_frozen = None

    launches = 0
    mode = {"now": "serial"}

    def start_exchange(bufs) -> int:
        runtime.check(lib.b200_event_record(ev_ready, main_stream))
        runtime.check(lib.b200_stream_wait_event(exchanger.stream, ev_ready))
        n = exchanger.exchange([(bufs["in_field"], HALO, HALO)])
        runtime.check(lib.b200_event_record(ev_halo, exchanger.stream))
        return n

    def step(i: int) -> int:
        """one pass of the hot path over this rank's slab"""
        bufs = sets[i & 1]
        if exchanger is None:
            return frozen(**bufs)
        if mode["now"] in ("peer", "peer_tma"):
            # push my boundary rows into the neighbours' halo (comm stream, behind everything enqueued so far on the compute
            # stream); the whole slab in one launch whose boundary tiles wait for the neighbours' flags on the device
            runtime.check(lib.b200_event_record(ev_ready, main_stream))
            runtime.check(lib.b200_stream_wait_event(peer.stream, ev_ready))
            n = peer.push([(bufs["in_field"], HALO, HALO)])
            return n + (peer_frozen if mode["now"] == "peer" else peer_tma_frozen)(**bufs, halo_wait=peer.wait_args())
        n = start_exchange(bufs)
        if mode["now"] == "serial":  # exchange, then the whole slab
            runtime.check(lib.b200_stream_wait_event(main_stream, ev_halo))
            return n + frozen(**bufs)
        if mode["now"] == "overlap":
            # halo exchange on the comm stream || interior rows on the compute stream; the two boundary strips
            # are whole tiles of the tuned kernel (not 2-row slivers) launched behind the interior
            n += frozen(**bufs, subbox=(0, NI, STRIP, NJ - STRIP))
            runtime.check(lib.b200_stream_wait_event(main_stream, ev_halo))
            n += frozen(**bufs, subbox=(0, NI, 0, STRIP))
            n += frozen(**bufs, subbox=(0, NI, NJ - STRIP, NJ))
            return n
        # "thin": interior on the compute stream; thin strips on the side stream behind the halo event only
        runtime.check(lib.b200_stream_wait_event(strip_stream, ev_ready))
        runtime.check(lib.b200_stream_wait_event(strip_stream, ev_halo))
        n += frozen(**bufs, subbox=(0, NI, THIN, NJ - THIN))
        n += thin_frozen(**bufs, subbox=(0, NI, 0, THIN), stream=strip_stream)
        n += thin_frozen(**bufs, subbox=(0, NI, NJ - THIN, NJ), stream=strip_stream)
        runtime.check(lib.b200_event_record(ev_strips, strip_stream))
        runtime.check(lib.b200_stream_wait_event(main_stream, ev_strips))
        return n

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if exchanger is not None and args.step_mode != "auto":
        # a forced schedule is also the one that is warmed up (first-call work — module load, lazy specialisation —
        # must not land in the timed region)
        avail_modes = ["serial", "overlap"] + (["thin"] if thin_frozen is not None else []) + (["peer"] if peer_frozen is not None else []) + (["peer_tma"] if peer_tma_frozen is not None else [])
        if args.step_mode not in avail_modes:
            raise SystemExit(f"bench.py: step mode {args.step_mode} is not available for this kernel ({peer_note})")
        mode["now"] = args.step_mode
    for i in range(args.warmup):
        step(i)
    barrier()
    overlap_trial = None
    if exchanger is not None:
        modes = ["serial", "overlap"] + (["thin"] if thin_frozen is not None else []) + (["peer"] if peer_frozen is not None else []) + (["peer_tma"] if peer_tma_frozen is not None else [])
        if args.step_mode != "auto":
            if args.step_mode not in modes:
                raise SystemExit(f"bench.py: step mode {args.step_mode} is not available for this kernel")
            mode["now"] = args.step_mode
        else:
            # every schedule must reproduce the serial one bit for bit on this rank's data before it may be timed
            expect, ok = None, {}
            for m in modes:
                mode["now"] = m
                sets[0]["out_field"].fill(0)
                step(0)
                torch.cuda.synchronize()
                got = sets[0]["out_field"].torch().clone()
                if expect is None:
                    expect = got
                ok[m] = bool(torch.equal(got, expect))
            flags = torch.tensor([1.0 if ok[m] else 0.0 for m in modes], device="cuda", dtype=torch.float64)
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)  # a schedule is usable only if it is exact on every rank
            usable = [m for m, f in zip(modes, flags.tolist()) if f > 0.5]
            # three rounds of 20 steps per schedule, interleaved; the median of the per-round maxima over ranks decides
            # (an 8-step trial was noise-level: r01 picked the serial schedule at N=4)
            rounds = {m: [] for m in usable}
            for _rep in range(3):
                for m in usable:
                    mode["now"] = m
                    for i in range(2):
                        step(i)
                    barrier()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for i in range(20):
                        step(i)
                    b.record()
                    barrier()
                    t = torch.tensor([a.elapsed_time(b) / 20], device="cuda", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    rounds[m].append(float(t.item()))
            trial = {m: round(sorted(v)[1], 5) for m, v in rounds.items()}
            mode["now"] = min(trial, key=trial.get)
            overlap_trial = {"ms_per_step": trial, "rejected_by_self_check": [m for m in modes if m not in usable]}

    pre_kernel_ms = None
    if exchanger is not None:  # N > 1: duration of the bare whole-slab kernel, before the timed region (same thermal state)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            frozen(**sets[i & 1])
        b.record()
        b.synchronize()
        pre_kernel_ms = a.elapsed_time(b) / 20
        barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream ----------------------
    # (clock sampling on rank 0 only: one nvidia-smi query stream for the whole job)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks is not None:
        clocks.__enter__()
    barrier()
    e0.record()
    for i in range(args.steps):
        launches += step(i)
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    # Keep the same load running long enough for the clock sampler to see it.  The number of extra
    # steps is derived from the ALL-REDUCED time, so every rank runs exactly the same count (the halo
    # exchange is a matched send/recv: a rank-dependent count would deadlock).
    extra_steps = int(min(4000, max(0, (700.0 - total_ms) / max(total_ms / args.steps, 1e-3))))
    for j in range(extra_steps):
        step(j)
    barrier()
    if clocks is not None:
        clocks.__exit__(None, None, None)
    ms_per_step = total_ms / args.steps
    cells_total = NI * NJ * NK * n_gpus
    value = cells_total / ms_per_step / 1e3  # Mcell/s

    # ---- dominant-kernel launch duration: back-to-back launches of the whole-domain stencil between
    # two events on the launching stream (no host gaps inside the region), average per launch
    kt = []
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            frozen(**sets[i & 1])
        b.record()
        b.synchronize()
        kt.append(a.elapsed_time(b) / 20)
    # At N = 1 a step IS one launch of the kernel: its average duration over the timed region is ms_per_step (same CUDA
    # events).  The back-to-back figure above comes after ~0.7 s of sustained load (clock sampling), i.e. under the
    # board's power cap (sw_power_cap, SM clock well below its maximum): reported next to it as the sustained duration.
    sustained_kernel_ms = float(np.mean(kt))
    kernel_ms = ms_per_step if n_gpus == 1 else pre_kernel_ms
    peak, peak_src = measured_peaks()
    achieved = NI * NJ * NK * BYTES_PER_CELL / kernel_ms / 1e6  # GB/s
    # DRAM traffic per launch: from the committed ncu capture, which is of the DEFAULT code-generation variant
    # (the tuned variants move the same rows through the same L2-prefetched streams; re-captured per round)
    traffic, traffic_src = None, None
    tr = ROOT / "profiles" / "hdiff_traffic.json"
    if tr.exists():
        try:
            # one ncu --set full capture per code-generation variant (tools/gpu_round.sh); the entry of the variant that
            # was timed if there is one, else the closest (most shared options) — the source string says which
            entries = json.loads(tr.read_text())
            entries = entries if isinstance(entries, list) else [entries]
            mine = {k: v for k, v in stencil.backend_options.items() if k not in ("strategy", "device_sync", "specialize")}
            def score(e):
                v = e.get("variant") or {}
                return (v == mine, sum(1 for k in v if mine.get(k) == v[k]) - sum(1 for k in set(v) ^ set(mine)))
            best = max(entries, key=score)
            traffic = best.get("dram_bytes_per_launch")
            traffic_src = f"{best.get('source')}; captured variant {best.get('variant')}" + ("" if best.get("variant") == mine else f" (timed variant: {mine})")
        except Exception:
            traffic = None

    # ---- N > 1: the decomposed result against ONE undecomposed domain, on the hardware ---------------------------------
    # One more step of the chosen schedule on buffer set 0; rank 0 gathers every rank's fields (NCCL), assembles the
    # global 1024 x (1024 N) x 80 problem, runs the single-GPU stencil on it and compares all rows of every slab bit for
    # bit (a wrong slab offset or a halo row that arrived late shows up at the slab boundaries).
    verify_note = None
    if dist is not None:
        try:
            sets[0]["out_field"].fill(0)
            step(0)
            barrier()
            loc = {n: sets[0][n].torch().contiguous() for n in ("in_field", "coeff", "out_field")}
            gathered = {}
            for n, t in loc.items():
                dst = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
                dist.gather(t, dst, dst=0)
                gathered[n] = dst
            flag = torch.zeros(1, device="cuda", dtype=torch.float64)
            if rank == 0:
                gshape = (shape[0], NJ * world + 2 * HALO, NK)
                glob = {}
                for n in ("in_field", "coeff"):
                    g = torch.empty(gshape, dtype=torch.float32, device="cuda")
                    g[:, :HALO] = gathered[n][0][:, :HALO]
                    g[:, HALO + NJ * world:] = gathered[n][world - 1][:, HALO + NJ:]
                    for r in range(world):
                        g[:, HALO + r * NJ: HALO + (r + 1) * NJ] = gathered[n][r][:, HALO: HALO + NJ]
                    glob[n] = storage.zeros(gshape, np.float32, aligned_index=origin3)
                    glob[n].torch().copy_(g)
                    del g
                glob["out_field"] = storage.zeros(gshape, np.float32, aligned_index=origin3)
                whole = B200Stencil(st_ir, {"strategy": args.strategy, "device_sync": True})  # default kernel, one domain
                whole(**glob, origin=origins, domain=(NI, NJ * world, NK))
                gout = glob["out_field"].torch()
                bad = 0
                for r in range(world):
                    want = gout[HALO: HALO + NI, HALO + r * NJ: HALO + (r + 1) * NJ]
                    got = gathered["out_field"][r][HALO: HALO + NI, HALO: HALO + NJ]
                    bad += int((want != got).sum().item())
                flag[0] = float(bad)
                del glob, gathered
            dist.broadcast(flag, src=0)
            verify_note = {"mismatching_cells_vs_single_domain": int(flag.item()), "cells": NI * NJ * NK * world, "schedule": mode["now"]}
        except Exception as exc:
            verify_note = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
        torch.cuda.empty_cache()

    # ---- the reference-facing plug-in: the SAME stencil built by gt4py's own frontend with backend="b200" -------------
    # (gt4py = the unmodified reference package from baseline/_ref; tools/install_reference.sh).  Used for the
    # end-to-end number below and for the per-call host cost of the public call paths.
    plug, plug_note = None, None
    try:
        if not _enable_reference():
            raise ImportError("gt4py (baseline/_ref) not importable")
        import warnings

        warnings.filterwarnings("ignore")
        import stencil_defs
        from gt4py.cartesian import gtscript

        import gt4py_b200

        if not gt4py_b200.register():  # registers backend="b200" and the storage hooks
            raise ImportError("gt4py_b200 plug-in did not register")
        case = stencil_defs.REGISTRY[STENCIL]
        popts = {k: v for k, v in stencil.backend_options.items() if k not in ("strategy", "device_sync")}
        plug = gtscript.stencil(backend="b200", definition=case["definition"], externals=case["externals"] or {},
                                name=f"{STENCIL}_bench_plug", device_sync=False, **case["build"], **popts)  # fmt: skip
    except Exception as exc:
        plug_note = f"plug-in path unavailable ({type(exc).__name__}: {str(exc)[:120]}): stand-alone mirror used"

    host_overhead = None
    if world == 1:
        # host time of one call (wall clock over 200 asynchronous calls; the kernel runs ~0.18 ms, so the queue never
        # drains): StencilObject.__call__ with and without argument validation, freeze(), CUDA-graph replay
        def host_us(fn, n=200):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            dt = time.perf_counter() - t0
            torch.cuda.synchronize()
            return round(dt / n * 1e6, 1)

        b0 = sets[0]
        host_overhead = {}
        try:
            target = plug if plug is not None else stencil
            kw = dict(in_field=b0["in_field"], out_field=b0["out_field"], coeff=b0["coeff"], origin=origins, domain=domain)
            target(**kw)  # first call: lazy loads
            host_overhead["StencilObject.__call__" if plug is not None else "B200Stencil.__call__"] = host_us(lambda: target(**kw))
            host_overhead["__call__(validate_args=False)"] = host_us(lambda: target(**kw, validate_args=False))
            fz = target.freeze(origin=origins, domain=domain)
            host_overhead["freeze()"] = host_us(lambda: fz(in_field=b0["in_field"], out_field=b0["out_field"], coeff=b0["coeff"]))
            from gt4py_b200.graph import StencilGraph

            g = StencilGraph()
            with g:
                frozen(**b0)
            host_overhead["StencilGraph.launch"] = host_us(g.launch)
            torch.cuda.synchronize()
            g.close()
        except Exception as exc:  # measurement code must not lose the bench line
            host_overhead["note"] = f"{type(exc).__name__}: {str(exc)[:160]}"

    def make_line(e2e):
        return {
            "metric": METRIC,
            "value": round(value, 1),
            "unit": "Mcell-updates/s",
            "n_gpus": n_gpus,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 5),
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "codegen_options": stencil.backend_options,
                "parallelism": "single GPU" if n_gpus == 1 else f"J-slab decomposition x{n_gpus}, NCCL halo exchange "
                + {"serial": "(exchange, then the whole slab)", "overlap": "overlapped with the interior tiles, whole-tile boundary strips behind",
                   "thin": f"overlapped with the interior, {THIN}-row boundary strips on a high-priority side stream",
                   "peer": "replaced by peer-memory pushes over NVLink (b200_halo_push into symmetric memory) consumed inside ONE stencil launch "
                           "per step (halo_wait kernel: boundary tiles last, device-side flag wait)",
                   "peer_tma": "replaced by peer-memory pushes over NVLink (b200_halo_push into symmetric memory) consumed inside ONE stencil launch "
                               "per step (halo_wait variant of the bulk-async kernel: boundary tiles last, device-side flag wait + cross-proxy fence)"}[mode["now"]],
                "schedule_trial": overlap_trial,
                "halo_wait_kernels": {"peer": hw_pick, "peer_tma": hwt_pick} if n_gpus > 1 else None,
                "multi_gpu_check": verify_note,
                "exposed_comm_us_per_step": round((ms_per_step - kernel_ms) * 1e3, 1) if n_gpus > 1 else None,
                "l2": "inputs larger than L2: 2 rotating buffer sets x 1.0 GB working set vs 126 MB L2",
                "kernels": stencil.compiled.kernel_names(),
                "host_us_per_call": host_overhead,
                "autotune_top5": tuned[:5] if isinstance(tuned, list) else tuned,
                "autotune_candidates": len(tuned) if isinstance(tuned, list) else 0,
                "autotune_rejected": getattr(stencil, "tune_rejected", None),
            },
            "gpu_launches": launches,
            "e2e": e2e,
            "roofline": {
                "bound": "hbm",
                "achieved": round(achieved, 1),
                "peak": peak,
                "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src,
                "kernel_ms": round(kernel_ms, 5),
                "kernel_ms_source": "the timed region itself (one launch per step)" if n_gpus == 1 else "20 back-to-back whole-slab launches before the timed region",
                "sustained_kernel_ms": round(sustained_kernel_ms, 5),
                "sustained_frac": round(NI * NJ * NK * BYTES_PER_CELL / sustained_kernel_ms / 1e6 / peak, 4),
                "algorithmic_bytes_per_launch": NI * NJ * NK * BYTES_PER_CELL,
            },
            "clocks": clocks.summary() if clocks is not None else None,
        }

    global _PARTIAL_LINE
    if rank == 0:  # what the watchdog prints if a later, optional phase hangs: the device-timed result is not lost
        _PARTIAL_LINE = make_line({"value": None, "unit": "Mcell-updates/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                   "note": "not measured: the run was aborted by the watchdog during the end-to-end phase"})
        if cpu_line is not None:
            _PARTIAL_LINE["cpu_baseline"] = cpu_line

    # ---- end-to-end through the public call with host buffers -------------------------------------
    # pinned host mirrors with the SAME pitched layout as the device storages, so every transfer is one
    # contiguous DMA of the padded buffer (not a strided element-wise copy over PCIe)
    from gt4py_b200 import hostpipe

    d_in, d_co, d_out = sets[0]["in_field"], sets[0]["coeff"], sets[0]["out_field"]
    stencil_e2e = B200Stencil(st_ir, {**stencil.backend_options, "device_sync": False})
    pin = {"in_field": hostpipe.PinnedMirror(d_in, host_in), "coeff": hostpipe.PinnedMirror(d_co, host_coeff),
           "out_field": hostpipe.PinnedMirror(d_out)}  # fmt: skip
    ti, tc, to = d_in._base, d_co._base, d_out._base
    nb = int(ti.numel()) * 4  # bytes actually transferred per field (padded pitch included)

    def e2e_step_serial():
        """copy in -> (halo exchange) -> public StencilObject-style call -> copy out, one stream"""
        ti.copy_(pin["in_field"].flat, non_blocking=True)
        tc.copy_(pin["coeff"].flat, non_blocking=True)
        if exchanger is not None:
            runtime.check(lib.b200_event_record(ev_ready, main_stream))
            runtime.check(lib.b200_stream_wait_event(exchanger.stream, ev_ready))
            exchanger.exchange([(d_in, HALO, HALO)])
            runtime.check(lib.b200_event_record(ev_halo, exchanger.stream))
            runtime.check(lib.b200_stream_wait_event(main_stream, ev_halo))
        stencil_e2e(d_in, d_out, d_co, origin=origins, domain=domain)
        pin["out_field"].flat.copy_(to, non_blocking=True)

    def time_e2e(fn, steps):
        for _ in range(2):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    e2e_steps = max(3, min(args.steps, 10))
    serial_ms = time_e2e(e2e_step_serial, e2e_steps)
    plug_ms = None
    if exchanger is None and not args.no_pipeline:
        # THE end-to-end path: StencilObject.__call__ of the plug-in (else of the mirror) with HOST arrays — pinned host
        # storages in the backend's layout (gt4py_b200.storage.host_from_array); the call stages them through device
        # storages (H2D / kernels / D2H of the compute domain, K-slab pipelined) and returns when the result is on the host
        try:
            target = plug if plug is not None else stencil_e2e
            h_in = storage.host_from_array(host_in, aligned_index=origin3)
            h_co = storage.host_from_array(host_coeff, aligned_index=origin3)
            h_out = storage.host_empty(shape, np.float32, aligned_index=origin3)
            info = {}
            target(h_in, h_out, h_co, origin=origins, domain=domain, exec_info=info)
            torch.cuda.synchronize()
            want = pin["out_field"].array[HALO:HALO + NI, HALO:HALO + NJ]
            if not np.array_equal(np.asarray(h_out)[HALO:HALO + NI, HALO:HALO + NJ], want):
                raise RuntimeError("host-array call differs from the device-resident call")
            plug_ms = time_e2e(lambda: target(h_in, h_out, h_co, origin=origins, domain=domain), e2e_steps)
            plug_path = (f"{'gt4py StencilObject.__call__ (backend=b200 plug-in)' if plug is not None else 'B200Stencil.__call__'} with pinned HOST "
                         f"storages: {info.get('b200_host_path')} (H2D in+coeff / kernels / D2H of the compute domain, K slabs on three streams)")
        except Exception as exc:
            plug_note = (plug_note + "; " if plug_note else "") + f"host-array call failed: {type(exc).__name__}: {str(exc)[:160]}"
    torch.cuda.synchronize()
    expect_out = pin["out_field"].flat.clone()  # result of the serial path (whole-domain kernel)
    e2e_ms, e2e_path = serial_ms, "serial: H2D(in, coeff) -> stencil call -> D2H(out) on one stream"
    pipe_note = None
    if exchanger is None and not args.no_pipeline:
        # the host-resident entry point of the backend: K-slab pipelining of H2D / stencil / D2H on three
        # streams (gt4py_b200/hostpipe.py); same bytes moved, both DMA directions and the SMs overlap
        try:
            pipe = hostpipe.HostPipeline(stencil_e2e, {"in_field": d_in, "coeff": d_co, "out_field": d_out},
                                         origin=origins, domain=domain, n_chunks=args.pipeline_chunks)  # fmt: skip
            pin["out_field"].flat.zero_()
            pipe(**pin)
            torch.cuda.synchronize()
            if not torch.equal(pin["out_field"].flat, expect_out):
                raise RuntimeError("pipelined result differs from the whole-domain call")
            pipe_ms = time_e2e(lambda: pipe(**pin), e2e_steps)
            if pipe_ms < serial_ms:
                e2e_ms = pipe_ms
                e2e_path = f"host pipeline: {len(pipe.chunks)} K slabs, H2D / stencil / D2H on three streams (bit-identical to the serial call)"
            else:
                pipe_note = f"host pipeline measured {pipe_ms:.3f} ms/step (slower than serial, not used)"
        except Exception as exc:  # measurement code must not lose the whole bench line
            pipe_note = f"host pipeline unavailable: {type(exc).__name__}: {exc}"
    d2h = nb
    if plug_ms is not None:
        # the headline end-to-end number is the public call itself; the lower-level HostPipeline / serial figures stay as context
        e2e_extra = {"hostpipe_value": round(cells_total / e2e_ms / 1e3, 1)}
        e2e_ms, e2e_path, d2h = plug_ms, plug_path, NI * NJ * NK * 4
    else:
        e2e_extra = {}
    e2e = {
        "value": round(cells_total / e2e_ms / 1e3, 1),
        "unit": "Mcell-updates/s",
        "h2d_bytes_per_step": 2 * nb,
        "d2h_bytes_per_step": d2h,
        "steps": e2e_steps,
        "path": e2e_path,
        "serial_value": round(cells_total / serial_ms / 1e3, 1),
        **e2e_extra,
    }
    if pipe_note or plug_note or peer_note:
        e2e["note"] = "; ".join(x for x in (pipe_note, plug_note, peer_note) if x)

    if rank == 0:
        line = make_line(e2e)
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        _PARTIAL_LINE = None  # the full line supersedes it (a hang at teardown must not print a second line)
        print(json.dumps(line), flush=True)
    if exchanger is not None:
        exchanger.close()
    if peer is not None:
        peer.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
