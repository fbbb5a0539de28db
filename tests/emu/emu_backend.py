"""TEST-ONLY gt4py backend "b200emu": the b200 code generator + the CPU emulator instead of the GPU.

Registered only by test tooling (never by the package), so that the REFERENCE's own test-suites
(test_code_generation.py, test_suites.py with hypothesis value checks) can exercise the b200 code
generators on machines without a GPU:   pytest -p emu.emu_backend_plugin ... -k b200emu
"""

from __future__ import annotations

import json
from typing import Any, ClassVar, Dict

from gt4py.cartesian import backend as gt_backend
from gt4py.storage.cartesian import layout as gt_layout

from gt4py_b200 import ir as b2ir
from gt4py_b200.backend import B200Backend, B200ModuleGenerator

from .emu import EmuStencil

_CACHE: Dict[str, EmuStencil] = {}


def run_emulated(ir_path: str, opts_json: str, domain, origin, fields, params) -> None:
    key = ir_path + "|" + opts_json
    es = _CACHE.get(key)
    if es is None:
        st = b2ir.load_file(ir_path)
        es = EmuStencil(st, json.loads(opts_json), name=st["name"])
        _CACHE[key] = es
    es.run({k: v for k, v in fields.items()}, params, tuple(int(d) for d in domain), origin)


class EmuModuleGenerator(B200ModuleGenerator):
    def generate_imports(self) -> str:
        return "import pathlib\nfrom gt4py.cartesian.stencil_object import StencilObject\nfrom emu import emu_backend as _emu"

    def generate_base_class_name(self) -> str:
        return "StencilObject"

    def generate_implementation(self) -> str:
        from gt4py.cartesian.gtc import gtir as gtir_mod

        gtir = self.builder.gtir
        fields = [p.name for p in gtir.params if isinstance(p, gtir_mod.FieldDecl)]
        params = [p.name for p in gtir.params if isinstance(p, gtir_mod.ScalarDecl)]
        fdict = ", ".join(f"{n}={n}" for n in fields)
        pdict = ", ".join(f"{n}={n}" for n in params)
        return f"_emu.run_emulated(_B200_IR, _B200_OPTS, _domain_, _origin_, dict({fdict}), dict({pdict}))"


@gt_backend.register
class B200EmuBackend(B200Backend):
    name = "b200emu"
    storage_info: ClassVar[gt_layout.LayoutInfo] = gt_layout.LayoutInfo(
        alignment=1,
        device="cpu",
        layout_map=gt_layout.layout_maker_factory((2, 1, 0)),
        is_optimal_layout=gt_layout.layout_checker_factory(gt_layout.layout_maker_factory((2, 1, 0))),
    )
    MODULE_GENERATOR_CLASS = EmuModuleGenerator
    compile_cubin = False  # nvcc is exercised by the real backend's tests; keep the emulated suites fast
