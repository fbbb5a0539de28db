"""Device comparison of the two staging designs of the streaming kernel (north_star: "TMA or shared-memory
staging ... the choices evidenced by ncu"): register-window loads (LDG + L2 prefetch) vs the bulk-async ring
(`tma`: cp.async.bulk -> shared-memory ring -> LDS).  Every variant must reproduce the default one bit for bit
before it is timed (B200Stencil.autotune).  Dev tool for one gpurun call:

    python tools/bench_tma.py [--workload hdiff|upwind5|pgrad|div|all] > gpurun_out/tma.jsonl
"""
import argparse
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "hdiff": ("hdiff_f32", "staged", (1024, 1024, 80)),
    "upwind5": ("upwind5_f32", "staged", (2048, 2048, 80)),
    "pgrad": ("fw_pgrad_f32", "staged", (4096, 512, 80)),
    "div": ("fw_div_f32", "staged", (4096, 512, 80)),
    "wsolve": ("fw_wsolve_f32", "default", (4096, 512, 80)),
    "tridiagonal": ("tridiagonal_f64", "default", (512, 512, 160)),
    "vadv": ("vadv_f64", "default", (512, 512, 160)),
}
P = {"interior_loop": True, "static_pitch": "auto"}
U = dict(P)  # (uniform_task is the default now; {"uniform_task": False} = the round-1 addressing)
T3 = {**P, "tma": 3, "tile_j": 32, "prefetch": 1}  # tensor-map copies (one per stream and trip) into a ring of 3 trips
CANDIDATES = [
    {}, dict(P), {**P, "min_blocks": 8}, {"interior_loop": "steady", "static_pitch": "auto", "row_pointers": True},
    {"uniform_task": False}, {**P, "uniform_task": False}, {**U, "min_blocks": 9}, {**U, "min_blocks": 10}, {**U, "min_blocks": 12}, {**U, "l2_prefetch": 4}, {**U, "l2_prefetch": 1},
    {**U, "prefetch": 2}, {**U, "prefetch": 0}, {**U, "prefetch": 2, "l2_prefetch": 4}, {**U, "vector_width": 4}, {**U, "vector_width": 4, "prefetch": 0},
    {**T3, "tma_mode": "bulk", "tma": 4}, {**T3, "tma_mode": "bulk", "tile_j": 16}, {**T3, "tma_mode": "bulk", "tile_j": 24}, {**T3, "tma_mode": "bulk", "tile_j": 48}, {**T3, "tma_mode": "bulk", "prefetch": 0}, {**U, "row_pointers": True},
    {**U, "tile_j": 32}, {**U, "tile_j": 128}, {**U, "warps": 2}, {**U, "warps": 8}, {**U, "stcs": True}, {**U, "warps": 8, "l2_prefetch": 4},
    dict(T3), {**T3, "tma": 4}, {**T3, "tile_j": 48}, {**T3, "tile_j": 24}, {**T3, "tile_j": 16}, {**T3, "tile_j": 64, "tma": 4}, {**T3, "tile_j": 64, "tma": 5, "warps": 2},
    {**T3, "warps": 2}, {**T3, "warps": 8, "tma_smem_kb": 48}, {**T3, "prefetch": 0}, {**T3, "prefetch": 0, "tma": 4}, {**T3, "tma_rows": 8, "tma_smem_kb": 48},
    {**T3, "tma_fence": False}, {**T3, "min_blocks": 9}, {**T3, "stcs": True}, {**T3, "l2_prefetch": 4}, {**T3, "tma": 3, "tma_mode": "bulk"},
]


def candidates_for(workload: str):
    """--candidates FILE: JSON {workload or "*": [options, ...]} replaces the built-in list (session-specific sweeps)."""
    for n, arg in enumerate(sys.argv):
        if arg == "--candidates":
            table = json.loads(pathlib.Path(sys.argv[n + 1]).read_text())
            return table.get(workload, table.get("*", []))
    return CANDIDATES


def precompile():
    """Build container: AOT-compile every candidate into the in-tree cubin cache (travels with gpurun)."""
    import math
    from concurrent.futures import ThreadPoolExecutor

    from gt4py_b200 import codegen, jit, testing

    jobs = []
    for wl, (name, variant, domain) in WORKLOADS.items():
        st = testing.load_ir(name, variant)
        shapes, _ = testing.field_layout(st, domain)
        widths = {math.ceil(s[0] / 32) * 32 for s in shapes.values() if len(s) == 3}
        for cand in candidates_for(wl):
            cand = dict(cand)
            if cand.get("static_pitch") == "auto":
                if len(widths) != 1:
                    continue
                cand["static_pitch"] = next(iter(widths))
            jobs.append((st, name, {"strategy": "auto", "device_sync": False, **cand}))

    def one(job):
        st, name, opts = job
        try:
            src, _plan = codegen.generate(st, opts)
            jit.compile_cubin(src, opts, name=codegen._cname(name))
        except Exception as exc:
            return f"{name} {opts}: {str(exc)[-300:]}"

    with ThreadPoolExecutor(8) as ex:
        for msg in ex.map(one, jobs):
            if msg:
                print("FAILED", msg)
    print(f"precompiled {len(jobs)} variants")


def main():
    if "--precompile" in sys.argv:
        return precompile()
    import numpy as np
    import torch

    from gt4py_b200 import storage, testing
    from gt4py_b200.stencil import B200Stencil

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hdiff")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--candidates", default=None)
    a = ap.parse_args()
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    for wl in WORKLOADS if a.workload == "all" else a.workload.split(","):
        name, variant, domain = WORKLOADS[wl]
        st = testing.load_ir(name, variant)
        fields, params, origins, domain = testing.make_case_data(st, name, domain=domain, seed=0)
        dev = {k: (storage.from_array(v, aligned_index=origins[k]) if v is not None else None) for k, v in fields.items()}
        s = B200Stencil(st, {"strategy": "auto", "device_sync": False})
        res = s.autotune(dev, params, domain=domain, origin=origins, iters=a.iters, candidates=candidates_for(wl), refine=0)
        bpc = testing.algorithmic_bytes_per_cell(st)
        cells = domain[0] * domain[1] * domain[2]
        for cand, ms in res:
            print(json.dumps({"workload": wl, "options": cand, "ms": round(ms, 5), "gbs": round(cells * bpc / ms / 1e6, 1),
                              "frac_of_peak": round(cells * bpc / ms / 1e6 / peak, 4)}), flush=True)
        print(json.dumps({"workload": wl, "rejected": getattr(s, "tune_rejected", None)}), flush=True)
        del dev
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
