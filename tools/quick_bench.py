"""Ad-hoc device timing of fixture stencils (dev tool; bench.py is the contract)."""
import argparse, json, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from gt4py_b200 import storage, testing
from gt4py_b200.stencil import B200Stencil

def bench(name, variant, strategy, domain, iters=20, warm=5, extra=None):
    st = testing.load_ir(name, variant)
    fields, params, origins, domain = testing.make_case_data(st, name, domain=domain, seed=0)
    dev = {k: (storage.from_array(v, aligned_index=origins[k]) if v is not None else None) for k, v in fields.items()}
    opts = {"strategy": strategy, "device_sync": False}; opts.update(extra or {})
    s = B200Stencil(st, opts)
    fr = s.freeze(origin=origins, domain=domain)
    kw = {**dev, **params}
    for _ in range(warm): fr(**kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):   # back-to-back launches between two events: no host-side gaps inside the timed region
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fr(**kw)
        e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1)/iters)
    ms = float(np.median(ts)); cells = domain[0]*domain[1]*domain[2]
    bpc = testing.algorithmic_bytes_per_cell(st)
    return {"name": name, "variant": variant, "strategy": strategy, "domain": domain, "ms": round(ms,4), "min_ms": round(min(ts),4),
            "mcells_s": round(cells/ms/1e3,1), "gbs": round(cells*bpc/ms/1e6,1), "launches": s.compiled.last_launches, **(extra or {})}

if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--cases", default="all"); ap.add_argument("--only", default=None); ap.add_argument("--name", default="hdiff_f32"); ap.add_argument("--variant", default="staged"); ap.add_argument("--domain", default="1024,1024,80"); ap.add_argument("--iters", type=int, default=20); ap.add_argument("--candidates", default=None, help="JSON list of option dicts: one line per variant of --name"); a = ap.parse_args()
    if a.candidates is not None:
        pk = json.loads((pathlib.Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
        for cand in json.loads(a.candidates):
            try:
                r = bench(a.name, a.variant, "auto", tuple(int(x) for x in a.domain.split(",")), iters=a.iters, extra=cand)
                r["frac_of_peak"] = round(r["gbs"] / pk, 4)
                print(json.dumps(r), flush=True)
            except Exception as e:
                print(json.dumps({"name": a.name, "options": cand, "error": repr(e)[:300]}), flush=True)
        sys.exit(0)
    if a.only is not None:
        print(json.dumps(bench(a.name, a.variant, "auto", tuple(int(x) for x in a.domain.split(",")), iters=a.iters, extra=json.loads(a.only)))); sys.exit(0)
    H=("hdiff_f32","staged","auto",(1024,1024,80))
    runs = [
        H, H+({"l2_prefetch":2},), H+({"l2_prefetch":2,"prefetch":0},),
        H+({"vector_width":2},), H+({"vector_width":2,"l2_prefetch":2},), H+({"vector_width":2,"l2_prefetch":4},),
        H+({"vector_width":2,"prefetch":0},), H+({"vector_width":2,"prefetch":0,"l2_prefetch":2},),
        H+({"vector_width":2,"prefetch":2},), H+({"vector_width":2,"prefetch":2,"l2_prefetch":4},),
        H+({"vector_width":2,"warps":2,"l2_prefetch":2},), H+({"vector_width":2,"warps":8,"l2_prefetch":2},),
        H+({"vector_width":2,"tile_j":64,"l2_prefetch":2},), H+({"vector_width":2,"tile_j":64,"prefetch":2},),
        H+({"tile_j":64,"l2_prefetch":2},), H+({"tile_j":64,"l2_prefetch":2,"warps":2},),
        H+({"vector_width":2,"min_blocks":10,"l2_prefetch":2},),
        ("copy_f64","default","auto",(1024,1024,40)),
        ("laplacian_f64","default","auto",(1024,1024,40)),
        ("upwind5_f32","staged","auto",(2048,2048,20),{"prefetch":0}),
        ("upwind5_f32","staged","auto",(2048,2048,20),{"prefetch":0,"vector_width":2}),
        ("tridiagonal_f64","default","point",(512,512,160)),
    ]
    for r in runs:
        try:
            extra = r[4] if len(r) > 4 else None
            print(json.dumps(bench(*r[:4], extra=extra)), flush=True)
        except Exception as e: print("FAIL", r, repr(e)[:500], flush=True)
