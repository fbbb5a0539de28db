"""Run ONE code-generation variant of a fixture stencil on the device against the oracle (dev tool; each variant in
its own process so that a device fault is attributable):  python tools/debug_variant.py hdiff_f32 '{"tma": 2, ...}' 256,128,4"""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import numpy as np
    import torch

    from gt4py_b200 import storage, testing
    from gt4py_b200.stencil import B200Stencil
    from oracle import numpy_oracle

    name, opts, dom = sys.argv[1], json.loads(sys.argv[2]), tuple(int(x) for x in sys.argv[3].split(","))
    st = testing.load_ir(name, "staged")
    fields, params, origins, dom = testing.make_case_data(st, name, domain=dom, seed=2)
    if opts.get("static_pitch") == "auto":
        import math

        shapes, _ = testing.field_layout(st, dom)
        opts["static_pitch"] = math.ceil(next(s[0] for s in shapes.values() if len(s) == 3) / 32) * 32
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, dom, origins)
    dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
    s = B200Stencil(st, {"strategy": "auto", **opts})
    s(**dev, **params, origin=origins, domain=dom)
    torch.cuda.synchronize()
    ok = all(np.array_equal(dev[f].get(), ref[f]) for f in testing.written_fields(st))
    print(json.dumps({"name": name, "options": opts, "domain": dom, "ok": bool(ok), "kernels": [(k["name"], k.get("tma"), k.get("tma_mode")) for k in s.compiled.plan["kernels"]]}))


if __name__ == "__main__":
    main()
