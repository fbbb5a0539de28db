"""Device timing of every BASELINE.json config (single-GPU part) over the code-generation variants
that still need a measurement.  Dev tool for one `gpurun` call; bench.py stays the contract.

    python tools/bench_configs.py [--quick] > gpurun_out/configs.jsonl

Each line: stencil, domain, options, ms per call (median of 5 x back-to-back launches between CUDA
events), Mcell/s, algorithmic GB/s (SURVEY §8d bytes) and its fraction of MEASURED_PEAKS.json.
"""
import argparse
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

from quick_bench import bench  # noqa: E402

PITCH = "auto"


def peak():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="quarter-size domains")
    a = ap.parse_args()
    q = 2 if a.quick else 1
    stream_variants = [
        {}, {"interior_loop": True}, {"interior_loop": True, "static_pitch": PITCH},
        {"interior_loop": True, "static_pitch": PITCH, "prefetch": 0},
        {"interior_loop": True, "static_pitch": PITCH, "vector_width": 4, "prefetch": 0},
        {"interior_loop": True, "static_pitch": PITCH, "l2_prefetch": 4},
    ]  # fmt: skip
    col_variants = [{"seq_cache": False}, {"seq_prefetch": 0}, {}, {"seq_prefetch": 2}, {"seq_prefetch": 3}, {"fuse_columns": True},
                    {"fuse_columns": True, "seq_prefetch": 2},
                    # fused sweeps with the resident columns capped so that a column's forward results are still in L2
                    # when its back substitution re-reads them (2 / 3 / 4 CTAs per SM), deeper look-ahead to make up
                    {"fuse_columns": True, "seq_prefetch": 4, "seq_smem_pad": 110 * 1024},
                    {"fuse_columns": True, "seq_prefetch": 4, "seq_smem_pad": 75 * 1024},
                    {"fuse_columns": True, "seq_prefetch": 6, "seq_smem_pad": 110 * 1024},
                    {"fuse_columns": True, "seq_prefetch": 3, "seq_smem_pad": 56 * 1024}]  # fmt: skip
    runs = []
    for v in stream_variants:
        runs.append(("hdiff_f32", "staged", (1024 // q, 1024 // q, 80), v))  # config 2
    for v in col_variants:
        runs.append(("tridiagonal_f64", "default", (512 // q, 512 // q, 160), v))  # config 3
        runs.append(("vadv_f64", "default", (512 // q, 512 // q, 160), v))
    for v in stream_variants:
        runs.append(("upwind5_f32", "staged", (2048 // q, 2048 // q, 80 // q), v))  # config 4 (one GPU)
    for name in ("fw_pgrad_f32", "fw_div_f32"):  # config 5, per-GPU share of 4096 x 4096 x 80 over 8 GPUs
        for v in stream_variants[:3]:
            runs.append((name, "staged", (4096 // q, 512 // q, 80), v))
    for v in col_variants[:3]:
        runs.append(("fw_wsolve_f32", "default", (4096 // q, 512 // q, 80), v))
    pk = peak()
    for name, variant, domain, extra in runs:
        extra = dict(extra)
        if extra.get("static_pitch") == "auto":
            import math

            from gt4py_b200 import testing

            st = testing.load_ir(name, variant)
            shapes, _ = testing.field_layout(st, domain)
            widths = {math.ceil(s[0] / 32) * 32 for s in shapes.values() if len(s) == 3}
            if len(widths) != 1:
                continue
            extra["static_pitch"] = widths.pop()
        try:
            r = bench(name, variant, "auto", domain, extra=extra)
            r["frac_of_peak"] = round(r["gbs"] / pk, 4)
            print(json.dumps(r), flush=True)
        except Exception as e:  # keep going: one failing variant must not lose the other measurements
            print(json.dumps({"name": name, "options": extra, "error": repr(e)[:300]}), flush=True)


if __name__ == "__main__":
    main()
