"""Benchmarks of the BASELINE.json configs other than the headline one (which is bench.py's):

  --workload tridiagonal   config 3: K-tridiagonal solve 512x512x160 fp64, FORWARD+BACKWARD sweeps (no halo; N>1 = replicas)
  --workload upwind5       config 4: 5th-order upwind advection 2048x2048x80 fp32, STRONG-scaled over N GPUs
                           (J slabs of 2048/N rows, NCCL halo exchange of phi, width 3, every step)
  --workload fastwaves     config 5: pressure-gradient -> divergence -> vertical implicit solve, 4096 x (512*N) x 80
                           fp32, WEAK-scaled (4096x4096x80 per box of 8); exchanges pp before the pressure
                           gradient and u, v before the divergence

Same launch / timing / JSON conventions as bench.py (torchrun for N>1, W warm-up steps, exactly K timed
steps between CUDA events, max over ranks, rank 0 prints one line with `roofline` and, at N=1,
`cpu_baseline`).  `--plan` prints the workload description without touching a device (CPU-testable).
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from gt4py_b200 import testing  # noqa: E402


def workload(name: str, n_gpus: int):
    """-> dict(stencils=[(fixture, variant)], domain (per rank), halo (I,J,K), exchanges={stencil index: [(field, width)]},
    scaling, label).  Fields of the same name are shared between the stencils of a step."""
    if name == "tridiagonal":
        return dict(stencils=[("tridiagonal_f64", "default")], domain=(512, 512, 160), halo=(0, 0, 0), exchanges={},
                    scaling="weak", dtype="f64", label="K-tridiagonal solve 512x512x160 fp64 per GPU (BASELINE configs[2]); replicas, no exchange")  # fmt: skip
    if name == "upwind5":
        if 2048 % n_gpus:
            raise SystemExit("upwind5: 2048 rows must divide by the number of GPUs")
        return dict(stencils=[("upwind5_f32", "staged")], domain=(2048, 2048 // n_gpus, 80), halo=(3, 3, 0),
                    exchanges={0: [("phi", 3)]}, scaling="strong", dtype="f32",
                    label=f"5th-order upwind advection 2048x2048x80 fp32 (BASELINE configs[3]), J slabs of {2048 // n_gpus} rows")  # fmt: skip
    if name == "fastwaves":
        return dict(stencils=[("fw_pgrad_f32", "staged"), ("fw_div_f32", "staged"), ("fw_wsolve_f32", "default")],
                    domain=(4096, 512, 80), halo=(1, 1, 1), exchanges={0: [("pp", 1)], 1: [("u", 1), ("v", 1)]},
                    rename={0: {"u_out": "u", "v_out": "v"}, 2: {"pp_out": "pp_new"}}, scaling="weak", dtype="f32",
                    label="fast-waves suite (pressure gradient, divergence, vertical implicit solve) 4096x512x80 fp32 per GPU "
                          "(BASELINE configs[4]: 4096x4096x80 per box of 8)")  # fmt: skip
    if name == "hdiff_x2":
        # two horizontal-diffusion time steps per pass (out of step 1 = in of step 2): as two calls (24 B per cell
        # and pass) or, with --fuse, as ONE fused stencil (gt4py_b200/fuse.py: 12 B per cell and pass, the hand-over
        # field lives in the register windows; step 1 is recomputed on a 2-cell rim of every tile)
        # N > 1 (weak-scaled J slabs): the two calls exchange a 2-row halo each (in_field, then mid); the fused
        # stencil exchanges ONE 4-row halo of in_field per pass (communication-avoiding: half the messages)
        return dict(stencils=[("hdiff_f32", "staged"), ("hdiff_f32", "staged")], domain=(1024, 1024, 80), halo=(4, 4, 0),
                    exchanges={0: [("in_field", 2)], 1: [("mid", 2)]}, fuse_exchanges={0: [("in_field", 4)]},
                    rename={0: {"out_field": "mid"}, 1: {"in_field": "mid"}}, fuse_intermediates=["mid"], updates_per_step=2,
                    scaling="weak", dtype="f32",
                    label="2 x horizontal diffusion 1024x1024x80 fp32 per GPU and pass (temporal blocking of BASELINE configs[1])")  # fmt: skip
    raise SystemExit(f"unknown workload {name}")


def fused_description(w, steps):
    """The whole step as ONE fused stencil."""
    from gt4py_b200 import fuse

    st = fuse.compose("_".join(dict.fromkeys(s["fixture"] for s in steps)) + f"_fused{len(steps)}",
                      [(s["ir"], s["binding"]) for s in steps], intermediates=w.get("fuse_intermediates", []))  # fmt: skip
    binding = {p["name"]: p["name"] for p in st["params"] if p["t"] == "field"}
    for fname, fi in st["field_info"].items():  # the fused stencil reads the combined halo: the workload must provide it
        if fi is not None and any(max(b) > w["halo"][a] for a, b in enumerate(fi["boundary"])):
            raise SystemExit(f"--fuse: field {fname} needs halo {fi['boundary']}, the workload allocates {w['halo']}")
    return [dict(fixture=st["name"], variant="fused", ir=st, binding=binding, bytes_per_cell=testing.algorithmic_bytes_per_cell(st),
                 params_from=[s["fixture"] for s in steps])]  # fmt: skip


def step_description(w):
    """Per stencil: IR, argument binding (stencil parameter -> shared buffer name), bytes per cell."""
    out = []
    for n, (fixture, variant) in enumerate(w["stencils"]):
        st = testing.load_ir(fixture, variant)
        ren = w.get("rename", {}).get(n, {})
        binding = {p["name"]: ren.get(p["name"], p["name"]) for p in st["params"] if p["t"] == "field"}
        out.append(dict(fixture=fixture, variant=variant, ir=st, binding=binding, bytes_per_cell=testing.algorithmic_bytes_per_cell(st)))
    return out


def plan(name: str, n_gpus: int):
    w = workload(name, n_gpus)
    steps = step_description(w)
    buffers = {}
    for s in steps:
        for p in s["ir"]["params"]:
            if p["t"] == "field":
                buffers.setdefault(s["binding"][p["name"]], p["dtype"])
    return dict(workload=name, n_gpus=n_gpus, domain_per_gpu=w["domain"], halo=w["halo"], scaling=w["scaling"],
                stencils=[s["fixture"] for s in steps], bytes_per_cell=[s["bytes_per_cell"] for s in steps],
                buffers=buffers, exchanges={str(k): v for k, v in w["exchanges"].items()}, label=w["label"])  # fmt: skip


def make_inputs(buffers, shape, rng):
    data = {}
    for name, dtype in buffers.items():
        a = rng.random(shape).astype(dtype)
        if name in ("diag", "rho"):
            a += 1.0
        if name in ("inf", "sup"):
            a *= 0.1
        if name == "hhl":  # monotone in K like a terrain-following height
            a += 100.0 * np.arange(shape[2], 0, -1, dtype=a.dtype)[None, None, :]
        data[name] = a.astype(dtype)
    return data


def default_params(st, fixture):
    spec = testing.CASE_SPECS.get(fixture, {}).get("params", {})
    out = {}
    for p in st["params"]:
        if p["t"] == "scalar" and st["parameter_info"].get(p["name"]) is not None:
            v = spec.get(p["name"], 0.75 if p["dtype"].startswith("float") else 2)
            out[p["name"]] = np.dtype(p["dtype"]).type(v)
    return out


def _reference_runners(steps):
    """The stencils of a step built by the REFERENCE numpy backend (gt4py from baseline/_ref, tools/refenv.py), or None
    when the reference is not importable here / a step is not a fixture definition (fused IRs)."""
    try:
        sys.path.insert(0, str(ROOT / "tools"))
        import refenv

        if not refenv.enable_gt4py():
            return None
        import warnings

        warnings.filterwarnings("ignore")
        import stencil_defs
        from gt4py.cartesian import gtscript

        runners = []
        for s in steps:
            case = stencil_defs.REGISTRY[s["fixture"]]
            runners.append(gtscript.stencil(backend="numpy", definition=case["definition"], externals=case["externals"] or {},
                                            name=f"{s['fixture']}_workload_ref", **case["build"]))  # fmt: skip
        return runners
    except Exception:
        return None


def cpu_baseline(steps, halo, sample=(128, 128, 16), updates=1):
    """CPU arm of a workload on a bounded sub-domain: the reference's own numpy backend (`kind: "reference"`) where gt4py is
    importable (baseline/_ref travels to the GPU box), else the oracle port."""
    rng = np.random.default_rng(0)
    shape = tuple(sample[a] + 2 * halo[a] for a in range(3))
    buffers = {}
    for s in steps:
        for p in s["ir"]["params"]:
            if p["t"] == "field":
                buffers.setdefault(s["binding"][p["name"]], p["dtype"])
    data = make_inputs(buffers, shape, rng)
    refs = _reference_runners(steps)
    if refs is not None:
        kind = "reference"

        def one():
            for s, ref in zip(steps, refs):
                fields = {p: data[b] for p, b in s["binding"].items()}
                ref(**fields, **default_params(s["ir"], s["fixture"]), origin={p: tuple(halo) for p in fields}, domain=tuple(sample))

    else:
        kind = "port"
        from oracle import numpy_oracle

        def one():
            for s in steps:
                fields = {p: data[b] for p, b in s["binding"].items()}
                numpy_oracle.run(s["ir"], fields, default_params(s["ir"], s["fixture"]), sample, {p: tuple(halo) for p in fields})

    try:
        one()
    except Exception:
        if kind != "reference":
            raise
        return cpu_baseline_port(steps, halo, sample, updates)
    times = []
    t_end = time.perf_counter() + 10.0
    while len(times) < 2 or (time.perf_counter() < t_end and len(times) < 30):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    cells = sample[0] * sample[1] * sample[2] * updates
    what = "reference gt4py numpy backend (baseline/_ref), StencilObject.__call__" if kind == "reference" else "oracle"
    return {"value": round(cells / float(np.median(times)) / 1e6, 3), "unit": "Mcell-updates/s", "cores": 1, "kind": kind,
            "sample": f"{what}, {sample[0]}x{sample[1]}x{sample[2]} sub-domain, {len(times)} steps, median ({os.cpu_count()} host cores available, NumPy uses 1)"}  # fmt: skip


def cpu_baseline_port(steps, halo, sample=(128, 128, 16), updates=1):
    """the oracle port, used when the reference is not importable or refuses the call"""
    prev = os.environ.get("B200_NO_REFERENCE")
    os.environ["B200_NO_REFERENCE"] = "1"
    try:
        return cpu_baseline(steps, halo, sample, updates)
    finally:
        if prev is None:
            os.environ.pop("B200_NO_REFERENCE", None)
        else:
            os.environ["B200_NO_REFERENCE"] = prev


P = {"interior_loop": True, "static_pitch": "auto"}
#: what --tune tries per stencil (a compact list: the winners of the round-2 sweeps, profiles/README.md)
TUNE_CANDIDATES = (
    {}, dict(P), {**P, "warps": 2}, {**P, "tile_j": 32}, {**P, "tile_j": 128}, {**P, "l2_prefetch": 1}, {**P, "prefetch": 0}, {**P, "prefetch": 1},
    {**P, "tma": 3, "tile_j": 32, "prefetch": 0}, {**P, "tma": 3, "tile_j": 32, "prefetch": 0, "tma_mode": "bulk"},
    {**P, "tma": 3, "tile_j": 32, "prefetch": 1, "tma_mode": "bulk"}, {**P, "tma": 3, "tile_j": 16, "prefetch": 1, "tma_mode": "bulk"},
    {**P, "tma": 3, "tile_j": 32, "prefetch": 1, "tma_mode": "bulk", "stcs": True},
    # column kernels (the streaming options above do not change their code: duplicates are skipped by the tuner)
    {"seq_rotate": False}, {"seq_rotate": True}, {"seq_prefetch": 2}, {"seq_prefetch": 4}, {"fuse_columns": True},
)  # fmt: skip


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["tridiagonal", "upwind5", "fastwaves", "hdiff_x2"])
    ap.add_argument("--fuse", action="store_true", help="run the stencils of a step as one fused stencil (gt4py_b200/fuse.py)")
    ap.add_argument("--peer", action="store_true", help="N>1: peer-memory halo exchange (symmetric memory pushes) instead of NCCL SendRecv")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--plan", action="store_true")
    ap.add_argument("--options", default="{}", help="JSON code-generation options for every stencil")
    ap.add_argument("--tune", action="store_true", help="N=1: autotune every stencil that can be re-run (B200Stencil.autotune over TUNE_CANDIDATES, "
                    "each candidate validated bit for bit against the default variant) before the timed steps")
    ap.add_argument("--shrink", type=int, default=1, help="divide the horizontal domain by this factor (smoke runs only)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.plan:
        print(json.dumps(plan(args.workload, args.gpus)))
        return

    import torch

    import bench as headline
    from gt4py_b200 import runtime, storage
    from gt4py_b200.distributed import HaloExchanger, PeerHalo, SlabDecomposition
    from gt4py_b200.stencil import B200Stencil

    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device — the b200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = workload(args.workload, world)
    steps = step_description(w)
    if args.fuse:
        if w["exchanges"] and world > 1 and "fuse_exchanges" not in w:
            raise SystemExit("--fuse: this workload exchanges halos between its stencils")
        steps = fused_description(w, steps)
        w["exchanges"] = w.get("fuse_exchanges", {})
        w["label"] += " [fused into one stencil]"
    ni, nj, nk = w["domain"]
    if args.shrink > 1:
        ni, nj = max(8, ni // args.shrink), max(8, nj // args.shrink)
        w["label"] += f" [SMOKE RUN: horizontal domain shrunk to {ni}x{nj}]"
    hi, hj, hk = w["halo"]
    shape = (ni + 2 * hi, nj + 2 * hj, nk + 2 * hk)
    origin3 = (hi, hj, hk)
    buffers = {}
    for s in steps:
        for p in s["ir"]["params"]:
            if p["t"] == "field":
                buffers.setdefault(s["binding"][p["name"]], p["dtype"])
    opts = {"device_sync": False, **json.loads(args.options)}
    # --peer (N > 1): exchanged fields in symmetric memory, halo rows pushed by the neighbours (b200_halo_push) and consumed
    # inside the stencil launch (halo_wait kernels) or behind a one-thread wait kernel — no NCCL call on the data path
    peer, exchanged = None, {f for fl in w["exchanges"].values() for f, _h in fl}
    if args.peer and world > 1 and w["exchanges"]:
        peer = PeerHalo(SlabDecomposition(world, rank, nj * world), nj)
    # two rotating buffer sets (working set >> L2)
    rng = np.random.default_rng(rank)
    sets = []
    for _ in range(2):
        host = make_inputs(buffers, shape, rng)
        sets.append({n: (peer.from_array(a, aligned_index=origin3) if (peer is not None and n in exchanged) else storage.from_array(a, aligned_index=origin3))
                     for n, a in sorted(host.items())})  # (sorted: symmetric allocations are collective, same order on every rank)
    frozen = []
    for k, s in enumerate(steps):
        sopts = dict(opts)
        if peer is not None and k in w["exchanges"]:
            sopts["halo_wait"] = True
        st = B200Stencil(s["ir"], sopts, name=f"{s['fixture']}.{s['variant']}")
        if args.tune and world == 1:
            try:
                tuned = st.autotune({p: sets[0][b] for p, b in s["binding"].items()}, default_params(s["ir"], s["fixture"]), domain=(ni, nj, nk),
                                    origin={p: origin3 for p in s["binding"]}, candidates=TUNE_CANDIDATES, iters=10, refine=2)  # fmt: skip
                s["tuned"] = {"options": tuned[0][0], "ms": tuned[0][1], "candidates": len(tuned)}
            except Exception as exc:  # the measurement must survive a tuner problem: default variant, reason recorded
                s["tuned"] = {"skipped": f"{type(exc).__name__}: {exc}"[:120]}
        frozen.append((st.freeze(origin={p: origin3 for p in s["binding"]}, domain=(ni, nj, nk)), default_params(s["ir"], s["fixture"]), st))
    exchanger = None
    lib = runtime.load_library()
    main_stream = torch.cuda.current_stream().cuda_stream
    if world > 1 and w["exchanges"]:
        import ctypes

        if peer is None:
            exchanger = HaloExchanger(SlabDecomposition(world, rank, nj * world), nj)
        ev = [ctypes.c_void_p(), ctypes.c_void_p()]
        for e in ev:
            runtime.check(lib.b200_event_create(ctypes.byref(e)))

    def step(i, exchange=True):
        bufs = sets[i & 1]
        n = 0
        for k, (s, (fr, params, _)) in enumerate(zip(steps, frozen)):
            if exchange and peer is not None and k in w["exchanges"]:
                runtime.check(lib.b200_event_record(ev[0], main_stream))
                runtime.check(lib.b200_stream_wait_event(peer.stream, ev[0]))
                n += peer.push([(bufs[f], hj, h) for f, h in w["exchanges"][k]])
                st = frozen[k][2]
                if all(kk["kind"] == "stream" for kk in st.compiled.plan["kernels"]):  # every kernel waits for itself
                    n += fr(**{p: bufs[b] for p, b in s["binding"].items()}, **params, halo_wait=peer.wait_args())
                else:
                    n += peer.wait(main_stream)
                    n += fr(**{p: bufs[b] for p, b in s["binding"].items()}, **params)
                continue
            if exchange and exchanger is not None and k in w["exchanges"]:
                runtime.check(lib.b200_event_record(ev[0], main_stream))
                runtime.check(lib.b200_stream_wait_event(exchanger.stream, ev[0]))
                n += exchanger.exchange([(bufs[f], hj, h) for f, h in w["exchanges"][k]])
                runtime.check(lib.b200_event_record(ev[1], exchanger.stream))
                runtime.check(lib.b200_stream_wait_event(main_stream, ev[1]))
            n += fr(**{p: bufs[b] for p, b in s["binding"].items()}, **params)
        return n

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = headline.ClockSampler(local_rank) if rank == 0 else None
    if clocks is not None:
        clocks.__enter__()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    launches = 0
    for i in range(args.steps):
        launches += step(i)
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    extra = int(min(2000, max(0, (600.0 - total_ms) / max(total_ms / args.steps, 1e-3))))
    for j in range(extra):  # same count on every rank (derived from the all-reduced time): matched exchanges
        step(j)
    barrier()
    if clocks is not None:
        clocks.__exit__(None, None, None)
    ms = total_ms / args.steps
    cells = ni * nj * nk * world * int(w.get("updates_per_step", 1))
    # kernels only, back to back (no exchange): the roofline numerator
    kt = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(10):
            step(i, exchange=False)
        b.record()
        b.synchronize()
        kt.append(a.elapsed_time(b) / 10)
    kernel_ms = float(np.mean(kt))
    bpc = sum(s["bytes_per_cell"] for s in steps)
    peak, peak_src = headline.measured_peaks()
    achieved = ni * nj * nk * bpc / kernel_ms / 1e6
    # DRAM traffic per step from the committed ncu captures of these kernels (profiles/workload_traffic.json: per kernel, the
    # variant that was captured is named there) — only for the domain the captures were taken on
    traffic, traffic_src = None, None
    tfile = ROOT / "profiles" / "workload_traffic.json"
    if tfile.exists() and args.shrink == 1:
        try:
            table = json.loads(tfile.read_text())
            names = [k for _, _, st in frozen for k in st.compiled.kernel_names()]
            have = [k for k in names if k in table and list(table[k]["domain"]) == [ni, nj, nk]]
            if have:
                traffic = float(sum(table[k]["dram_bytes_per_launch"] for k in have))
                missing = [k for k in names if k not in have]
                traffic_src = ("profiles/workload_traffic.json: sum over " + ", ".join(f"{k} (captured variant {table[k]['captured_variant']})" for k in have)
                               + (f"; no capture for {', '.join(missing)}" if missing else ""))
        except Exception:
            traffic, traffic_src = None, None
    if rank == 0:
        line = {
            "metric": f"Mcell-updates/s + achieved HBM GB/s, {args.workload}",
            "value": round(cells / ms / 1e3, 1), "unit": "Mcell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms, 5), "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": w["dtype"],
            "data": "synthetic",
            "config": {"workload": w["label"], "domain_per_gpu": [ni, nj, nk], "halo": list(w["halo"]),
                       "stencils": [s["fixture"] for s in steps], "codegen_options": opts, "tuned": [s.get("tuned") for s in steps] if args.tune else None,
                       "exchange": None if world == 1 or not w["exchanges"] else ("peer-memory pushes (b200_halo_push, symmetric memory) consumed by halo_wait kernels" if peer is not None else "NCCL SendRecv (b200_halo_exchange), serial"),
                       "kernels": [st.compiled.kernel_names() for _, _, st in frozen],
                       "l2": "two rotating buffer sets, each larger than the 126 MB L2"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms": round(kernel_ms, 5),
                         "algorithmic_bytes_per_launch": ni * nj * nk * bpc, "bytes_per_cell": bpc},
            "clocks": clocks.summary() if clocks is not None else None,
        }  # fmt: skip
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(steps, w["halo"], updates=int(w.get("updates_per_step", 1)))
        print(json.dumps(line))
    if exchanger is not None:
        exchanger.close()
    if peer is not None:
        peer.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
