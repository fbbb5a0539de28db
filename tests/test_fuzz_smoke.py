"""Keeps tools/fuzz_codegen.py (differential fuzzing: reference numpy backend == oracle == emulated
generated kernels) alive: one seeded random stencil per family.  The real campaigns are run by hand
(`python tools/fuzz_codegen.py --n 200`); their results are recorded in DESIGN.md §9."""

import pathlib
import random
import sys

import pytest

pytestmark = pytest.mark.needs_gt4py
ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("family,seed", [("par", 300007), ("col", 400003)])
def test_one_random_stencil(family, seed, tmp_path):
    sys.path.insert(0, str(ROOT / "tools"))
    import fuzz_codegen as fz

    rng = random.Random(seed)
    dtype = rng.choice(["float32", "float64"])
    gen = fz.Gen(rng, dtype)
    name = f"fzt_{family}_{seed}"
    source = gen.par(name) if family == "par" else gen.col(name)
    res = fz.run_case(source, name, dtype, seed, tmp_path, {})
    assert res == "" or res.startswith("SKIP"), res + "\n" + source
