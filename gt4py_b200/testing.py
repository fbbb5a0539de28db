"""Seeded synthetic inputs for the fixture stencils (shared by the golden generator, tests, bench).

No gt4py dependency: everything is derived from the lowered IR (`tests/golden/ir/*.json`).
"""

from __future__ import annotations

import pathlib
from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import ir as b2ir

GOLDEN_DIR = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"
IR_DIR = GOLDEN_DIR / "ir"

DEFAULT_DOMAIN = (20, 13, 8)

#: per-case overrides: domain, scalar parameters, special input recipes
CASE_SPECS: Dict[str, Dict[str, Any]] = {
    "laplacian_f64": {"domain": (64, 64, 16)},  # BASELINE.json configs[0]
    "hdiff_f32": {"domain": (40, 21, 6)},
    "hdiff_f32_default_literals": {"domain": (40, 21, 6)},
    "tridiagonal_f64": {"inputs": "tridiagonal"},
    "vadv_f64": {"params": {"dtr_stage": 3.0 / 20.0}},
    "upwind5_f32": {
        "domain": (33, 18, 5),
        "params": {"dt": np.float32(0.01), "dx": np.float32(0.1), "dy": np.float32(0.125)},
    },
    "fw_pgrad_f32": {"inputs": "fastwaves", "params": {"dt": np.float32(0.5), "edadlat": np.float32(0.01)}},
    "fw_div_f32": {"inputs": "fastwaves", "params": {"edadlat": np.float32(0.01)}},
    "fw_wsolve_f32": {"inputs": "fastwaves", "params": {"dt": np.float32(0.5), "c2": np.float32(0.3)}},
    "scale_param_f32": {"params": {"alpha": np.float32(1.75), "n": np.int32(3)}},
    "if_scalar_f64": {"params": {"flag": np.int32(1)}},
    "varoff_f64": {"inputs": "varoff"},
    "col_backward_f64": {"inputs": "varoff_up"},
    "col_multiwrite_f32": {"params": {"w": np.float32(0.625)}},
    "ints_bools": {"inputs": "ints"},
    "casts": {"inputs": "ints"},
    "sections_koff_f64": {"domain": (70, 40, 6)},
    "div_param_f32": {"inputs": "div_specials", "domain": (70, 23, 3), "params": {"dx": np.float32(0.1), "dy": np.float32(-2.5e-3)}},
    "div_param_col_f64": {"inputs": "div_specials", "domain": (37, 9, 7), "params": {"dz": np.float64(0.3)}},
}


def list_cases():
    return sorted({p.name.split(".")[0] for p in IR_DIR.glob("*.json")})


def load_ir(name: str, variant: str = "default") -> Dict[str, Any]:
    return b2ir.load_file(IR_DIR / f"{name}.{variant}.json")


def _np_dtype(name: str):
    return np.dtype("bool" if name == "bool" else name)


def field_layout(stencil, domain, extra_halo=(1, 0, 1)) -> Tuple[Dict[str, Tuple[int, ...]], Dict[str, Tuple[int, ...]]]:
    """Shapes and origins of the API fields for `domain` (halo = boundary + a little extra)."""
    shapes, origins = {}, {}
    for p in stencil["params"]:
        if p["t"] != "field":
            continue
        fi = stencil["field_info"].get(p["name"])
        if fi is None:
            continue
        shape, origin = [], []
        for ax, present in zip("IJK", p["dims"]):
            if not present:
                continue
            a = "IJK".index(ax)
            lo, hi = fi["boundary"][a]
            lo += extra_halo[a]
            shape.append(lo + domain[a] + hi + (1 if a == 1 else 0))
            origin.append(lo)
        shape += list(p["data_dims"])
        origin += [0] * len(p["data_dims"])
        shapes[p["name"]] = tuple(shape)
        origins[p["name"]] = tuple(origin)
    return shapes, origins


def make_case_data(stencil, name: Optional[str] = None, domain=None, seed: int = 0):
    """-> (fields: name->np.ndarray (C order, IJK[+data] axes), params, origins, domain)."""
    name = name or stencil["name"]
    spec = CASE_SPECS.get(name, {})
    domain = tuple(domain or spec.get("domain", DEFAULT_DOMAIN))
    rng = np.random.default_rng(seed)
    shapes, origins = field_layout(stencil, domain)
    recipe = spec.get("inputs")
    fields: Dict[str, Any] = {}
    for p in stencil["params"]:
        if p["t"] != "field":
            continue
        if p["name"] not in shapes:
            fields[p["name"]] = None
            continue
        dt = _np_dtype(p["dtype"])
        shape = shapes[p["name"]]
        if dt.kind == "f":
            arr = rng.random(shape).astype(dt)
        elif dt.kind == "b":
            arr = rng.random(shape) > 0.5
        else:
            arr = rng.integers(-10, 10, size=shape).astype(dt)
        fields[p["name"]] = arr
    if recipe == "tridiagonal":
        fields["inf"] *= 0.1
        fields["sup"] *= 0.1
        fields["diag"] += 1.0
    elif recipe == "varoff":
        fields["idx"] = rng.integers(-2, 3, size=shapes["idx"]).astype(np.int32)
    elif recipe == "varoff_up":
        fields["idx"] = rng.integers(0, 3, size=shapes["idx"]).astype(np.int32)
    elif recipe == "div_specials":
        # operands that leave the guarded exponent range of the hoisted-reciprocal division (csrc/b200_device.cuh, DivInv)
        for n, arr in fields.items():
            if arr is None or arr.dtype.kind != "f" or n.startswith("out"):
                continue
            tiny, huge = (1e-44, 1e30) if arr.dtype == np.float32 else (5e-324, 1e300)
            flat = arr.reshape(-1)
            idx = rng.permutation(flat.size)
            m = max(1, flat.size // 16)
            for j, val in enumerate((0.0, -0.0, np.inf, -np.inf, np.nan, tiny, -tiny)):
                flat[idx[j * m:(j + 1) * m]] = val
            flat[idx[7 * m:8 * m]] *= arr.dtype.type(huge)
            flat[idx[8 * m:9 * m]] *= arr.dtype.type(1e-30 if arr.dtype == np.float32 else 1e-290)
            flat[idx[9 * m:10 * m]] -= arr.dtype.type(0.5)
    elif recipe == "fastwaves":
        if "hhl" in fields and fields["hhl"] is not None:
            nk = fields["hhl"].shape[2]
            fields["hhl"] = (fields["hhl"] + 100.0 * np.arange(nk, 0, -1, dtype=np.float32)[None, None, :]).astype(np.float32)
        if "rho" in fields and fields["rho"] is not None:
            fields["rho"] = (fields["rho"] + 1.0).astype(np.float32)
    params = {}
    for p in stencil["params"]:
        if p["t"] == "scalar":
            pi = stencil["parameter_info"].get(p["name"])
            if pi is None:
                continue
            val = spec.get("params", {}).get(p["name"])
            if val is None:
                val = 0.75 if p["dtype"].startswith("float") else 2
            params[p["name"]] = _np_dtype(p["dtype"]).type(val)
    return fields, params, origins, domain


def written_fields(stencil):
    return [n for n, fi in stencil["field_info"].items() if fi is not None and fi["access"] in ("WRITE", "READ_WRITE")]


def algorithmic_bytes_per_cell(stencil) -> int:
    """sum over API fields of itemsize x (1 for R or W, 2 for RW), 3-D fields only count fully."""
    total = 0
    for name, fi in stencil["field_info"].items():
        if fi is None or fi["access"] == "NONE":
            continue
        n = b2ir.ITEMSIZE["bool" if fi["dtype"] == "bool" else fi["dtype"]]
        for d in fi["data_dims"]:
            n *= d
        if len(fi["axes"]) < 3:
            continue  # lower-dimensional fields are O(surface), not counted per cell
        total += n * (2 if fi["access"] == "READ_WRITE" else 1)
    return total
