"""CPU execution of the GENERATED CUDA kernels (tests/emu: g++ + a warp-lockstep shim) against the
oracle and the reference goldens.  Validates both code generators without a GPU; the GPU suite
(`-m gpu`) repeats the same comparisons on the real device through the C-ABI launcher."""

import numpy as np
import pytest

from gt4py_b200 import testing
from oracle import numpy_oracle

from emu.emu import EmuStencil
from parity_util import compare

CASES = testing.list_cases()


def run_emulated(name, variant, options, domain=None, seed=0, check_golden=False, subboxes=None, layout=None, guard=None):
    st = testing.load_ir(name, variant)
    if domain is not None:
        domain = (domain[0], domain[1], max(domain[2], int(st["domain_info"]["min_k"])))
    fields, params, origins, domain = testing.make_case_data(st, name, domain=domain, seed=seed)
    ref = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    es = EmuStencil(st, options, name=f"{name}.{variant}")
    es.trace()
    for box in subboxes or [None]:
        es.run(fields, params, domain, origins, subbox=box, layout=layout, guard=guard)
    for fname, fi in st["field_info"].items():
        if fi is not None and fields.get(fname) is not None:
            compare(name, fname, fields[fname], ref[fname])
    if check_golden:
        golden = np.load(testing.GOLDEN_DIR / f"{name}.npz")
        for fname in testing.written_fields(st):
            compare(name, fname, fields[fname], golden[f"seed{seed}.{fname}"])
    return es


@pytest.mark.parametrize("name", CASES)
def test_default_strategy_on_staged_lowering(name):
    """what `backend="b200"` generates (streaming kernels where applicable)"""
    run_emulated(name, "staged", {"strategy": "auto"}, check_golden=True)


@pytest.mark.parametrize("name", CASES)
def test_default_strategy_in_the_storage_layout(name):
    """same, with the fields in the backend's storage layout (I unit-stride, padded, aligned origin):
    the layout gt4py.storage hands out, which selects the 16-byte vector path and the steady-state
    loop of the streaming kernels; buffers end / start against guard pages (no out-of-bounds access)"""
    run_emulated(name, "staged", {"strategy": "auto"}, check_golden=True, layout="b200", guard="end")
    run_emulated(name, "default", {"strategy": "auto"}, seed=1, check_golden=True, layout="b200", guard="start")


def test_steady_loops_are_what_the_storage_layout_runs():
    """path coverage: with gt4py.storage-style fields most march steps of horizontal diffusion
    go through the (pure) steady loop; with odd extents the edge warps take the edge loop when it is
    generated and the general loop otherwise — bit-identical results either way"""
    for opts, steady_slot in (({}, 0), ({"edge_loop": True}, 0), ({"pure_loop": False}, 1)):
        es = run_emulated("hdiff_f32", "staged", opts, domain=(151, 70, 2), seed=12, layout="b200", guard="end")
        tr = es.trace()
        # (one of the three I segments of this small domain holds the odd last column -> an edge warp)
        assert tr[steady_slot] > 0 and tr[steady_slot] * es.plan["kernels"][0]["period"] > tr[2], (opts, tr)
        if opts.get("edge_loop"):
            assert tr[1] > 0, tr
    es = run_emulated("hdiff_f32", "staged", {}, domain=(151, 70, 2), seed=12)  # C-order arrays: general loop only
    assert es.trace()[0] == 0
    # interior loop (no per-lane predicates) + compile-time row pitch: taken by the full-width warps
    es = run_emulated("hdiff_f32", "staged", {"interior_loop": True, "static_pitch": 160}, domain=(151, 70, 2), seed=12, layout="b200", guard="end")
    tr = es.trace()
    assert tr[3] > 0 and tr[0] > 0, tr
    es = run_emulated("hdiff_f32", "staged", {"interior_loop": True, "static_pitch": 192}, domain=(151, 70, 2), seed=12, layout="b200", guard="end")
    tr = es.trace()
    assert tr[3] == 0 and tr[0] == 0 and tr[2] > 0, tr


@pytest.mark.parametrize("name", CASES)
def test_point_generator_on_default_lowering(name):
    run_emulated(name, "default", {"strategy": "point"}, seed=1, check_golden=True)


@pytest.mark.parametrize("name", ["hdiff_f32", "upwind5_f32", "laplacian_f64", "two_stage_par_f32", "fw_pgrad_f32", "regions_f64",
                                  "fuse_chain_f32", "fuse_reuse_f64", "fuse_partial_f64", "stage_halo_f32"])  # fmt: skip
@pytest.mark.parametrize("domain", [(1, 1, 1), (3, 2, 1), (61, 5, 2), (129, 70, 2)])
def test_ragged_domains_streaming(name, domain):
    run_emulated(name, "staged", {"strategy": "auto"}, domain=domain, seed=2)
    run_emulated(name, "staged", {"strategy": "auto"}, domain=domain, seed=3, layout="b200", guard="end")
    run_emulated(name, "staged", {"strategy": "auto"}, domain=domain, seed=3, layout="b200", guard="start")


@pytest.mark.parametrize("name,variant", [("hdiff_f32", "staged"), ("fw_pgrad_f32", "staged"), ("upwind5_f32", "staged")])
@pytest.mark.parametrize("strategy", ["auto", "point"])
def test_subbox_launches_compose(name, variant, strategy):
    ni, nj = 70, 150
    boxes = [(0, ni, 64, nj - 64), (0, ni, 0, 64), (0, ni, nj - 64, nj)]
    run_emulated(name, variant, {"strategy": strategy}, domain=(ni, nj, 2), seed=3, subboxes=boxes)
    boxes = [(0, 37, 0, nj), (37, 61, 0, 3), (37, 61, 3, nj), (61, ni, 0, nj)]
    run_emulated(name, variant, {"strategy": strategy}, domain=(ni, nj, 2), seed=4, subboxes=boxes)
    run_emulated(name, variant, {"strategy": strategy}, domain=(ni, nj, 2), seed=4, subboxes=boxes, layout="b200", guard="end")
    if strategy == "auto":  # the tuned variants bench.py may pick, launched strip by strip (multi-GPU overlap)
        tuned = {"interior_loop": True, "static_pitch": 96, "tile_j": 32}
        boxes = [(0, ni, 32, nj - 32), (0, ni, 0, 32), (0, ni, nj - 32, nj)]
        run_emulated(name, variant, tuned, domain=(ni, nj, 2), seed=5, subboxes=boxes, layout="b200", guard="end")


@pytest.mark.parametrize("tuned", [{}, {"interior_loop": True, "static_pitch": 96}, {"vector_width": 4}])
def test_thin_strip_schedule_composes_two_kernel_variants(tuned):
    """bench.py's "thin" multi-GPU schedule: interior rows by the tuned kernel, the two 16-row boundary strips
    by the tile_j=16 variant of the same kernel (side stream on the device) == the whole-domain result"""
    st = testing.load_ir("hdiff_f32", "staged")
    ni, nj, thin = 70, 150, 16
    fields, params, origins, domain = testing.make_case_data(st, "hdiff_f32", domain=(ni, nj, 2), seed=8)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    EmuStencil(st, {"strategy": "auto", **tuned}, name="hdiff_f32.staged").run(
        fields, params, domain, origins, subbox=(0, ni, thin, nj - thin), layout="b200", guard="end")
    strips = EmuStencil(st, {"strategy": "auto", **tuned, "tile_j": thin}, name="hdiff_f32.staged")
    assert strips.plan["kernels"][0]["tile"][1] == thin
    for box in ((0, ni, 0, thin), (0, ni, nj - thin, nj)):
        strips.run(fields, params, domain, origins, subbox=box, layout="b200", guard="end")
    compare("hdiff_f32", "out_field", fields["out_field"], ref["out_field"])


@pytest.mark.parametrize(
    "opts",
    [
        {"vector_width": 4}, {"vector_width": 2, "tile_j": 16, "warps": 2}, {"prefetch": 0, "l2_prefetch": 0},
        {"prefetch": 2, "tile_j": 8}, {"edge_loop": True}, {"pure_loop": False}, {"vector_width": 4, "edge_loop": True},
        {"tile_j": 128}, {"l2_prefetch": 4}, {"warps": 8},
        {"interior_loop": True}, {"static_pitch": 160}, {"interior_loop": True, "static_pitch": 160},
        {"interior_loop": True, "static_pitch": 160, "vector_width": 4, "prefetch": 0},
        {"interior_loop": True, "static_pitch": 160, "l2_prefetch": 0, "tile_j": 32, "warps": 2},
        {"interior_loop": True, "static_pitch": 160, "stcs": True, "ldcs": True}, {"stcs": True},
        {"interior_loop": True, "static_pitch": 160, "row_pointers": True}, {"static_pitch": 160, "row_pointers": True},
        {"interior_loop": "steady", "static_pitch": 160}, {"interior_loop": True, "static_pitch": 160, "tile_j": 52, "min_blocks": 8},
        {"interior_loop": True, "static_pitch": 192},  # pitch of the arguments differs -> general loop, same result
        {"k_order": True}, {"k_order": False, "period": 8, "div_slow": "call"}, {"period": 4, "div_slow": "inline"}, {"div_inv": False},
        {"interior_loop": True, "static_pitch": 160, "period": 8}, {"interior_loop": True, "static_pitch": 160, "k_order": True, "halo_wait": True},
    ],
)  # fmt: skip
def test_streaming_tuning_knobs_do_not_change_results(opts):
    for name in ("hdiff_f32", "upwind5_f32"):
        run_emulated(name, "staged", {"strategy": "auto", **opts}, domain=(75, 41, 2), seed=5)
        run_emulated(name, "staged", {"strategy": "auto", **opts}, domain=(139, 70, 2), seed=5, layout="b200", guard="end")


SEQ_CASES = ["tridiagonal_f64", "vadv_f64", "fw_wsolve_f32", "fwd_scan_f64", "lowdim_write_f64", "col_mask_f64",
             "col_chain_f64", "col_backward_f64", "col_multiwrite_f32"]  # fmt: skip


@pytest.mark.parametrize("name", SEQ_CASES)
@pytest.mark.parametrize("opts", [{}, {"seq_prefetch": False}, {"seq_prefetch": 2}, {"seq_prefetch": 3}, {"seq_cache": False},
                                  {"fuse_columns": True}, {"fuse_columns": True, "seq_prefetch": 2},
                                  {"col_smem": True}, {"col_smem": True, "seq_prefetch": 4}, {"col_smem": True, "col_smem_kb": 1}, {"col_smem": True, "col_smem_kb": 3}, {"seq_rotate": False, "seq_prefetch": 3},
                                  {"fuse_columns": True, "seq_prefetch": 5, "col_hints": True}, {"col_hints": True, "seq_prefetch": 2},
                                  {"col_smem": True, "col_smem_block": (64, 2), "fuse_columns": True, "div_inv": False}])  # fmt: skip
def test_column_generator_variants(name, opts):
    """register k-cache column kernels (default), without the one-level-ahead prefetch, and the
    baseline column kernel: same results on ragged domains, both lowerings"""
    for variant, domain, seed in (("default", (37, 5, 9), 6), ("staged", (3, 2, 4), 7)):
        es = run_emulated(name, variant, opts, domain=domain, seed=seed, layout="b200" if seed == 6 else None, guard="end")
        kinds = {k["name"].rsplit("_", 1)[-1][:3] for k in es.plan["kernels"]}
        assert ("col" in kinds) == (opts.get("seq_cache", True)), kinds


def test_column_generator_keeps_the_k_dataflow_in_registers():
    """Thomas forward sweep: per level 4 prefetched loads (inf, diag, sup, rhs), 2 stores, and the
    k-1 values of sup / rhs carried in registers (no load)."""
    from gt4py_b200 import codegen

    src, plan = codegen.generate(testing.load_ir("tridiagonal_f64", "default"), {"fuse_columns": True})
    assert [k["name"] for k in plan["kernels"]] == ["b200_tridiagonal_f64_col0"]  # both sweeps, one launch
    sweep = src[src.index("forward sweep 0, section 1"):src.index("backward sweep 1, section 0")]
    assert len(codegen.generate(testing.load_ir("tridiagonal_f64", "default"), {})[1]["kernels"]) == 2  # default: one per sweep
    loop = sweep[sweep.index("for (int k"):]
    assert "carried ['sup_p0p0m1', 'rhs_p0p0m1']" in sweep
    # the forward elimination updates sup / rhs in place: shifting pipeline, one level of look-ahead (see codegen_column.py)
    assert "// ring slot" not in loop and "n_inf_p0p0p0" in loop
    assert loop.count("b200::ldro<double>") == 2 and loop.count("= c_sup[") == 1 and loop.count("= c_rhs[") == 1
    assert loop.count("c_sup[(long long)k *") == 1 and loop.count("c_rhs[(long long)k *") == 1  # one store each
    # the back substitution only reads what it looks ahead for: a ring of 3 + 1 registers per value, the level loop
    # unrolled four times, no moves between the slots
    back = src[src.index("backward sweep 1, section 1"):]
    assert back.count("// ring slot") == 4 and "q3_sup_p0p0p0 = c_sup[" in back and "= q1_sup" not in back.replace("* q1_sup", "")
    ring = codegen.generate(testing.load_ir("tridiagonal_f64", "default"), {"fuse_columns": True, "seq_rotate": True, "seq_prefetch": 1})[0]
    fwd = ring[ring.index("forward sweep 0, section 1"):ring.index("backward sweep 1, section 0")]
    assert fwd.count("// ring slot") == 2 and "q1_inf_p0p0p0 = b200::ldro<double>" in fwd and "= q1_inf" not in fwd
    assert "(k + (-1))" not in loop


@pytest.mark.parametrize("name", ["fw_pgrad_f32", "fuse_chain_f32", "fuse_reuse_f64", "fuse_partial_f64", "tmp_koffset_f64", "k_intervals_f64"])
def test_loop_fusion_by_interval_refinement(name):
    """consecutive PARALLEL computations fused into one set of kernels (one per K interval piece) give
    the results of the unfused stencil; values handed between them no longer go through scratch"""
    from gt4py_b200 import codegen

    st = testing.load_ir(name, "staged")
    fused, plain = codegen.generate(st, {})[1], codegen.generate(st, {"fuse_loops": False})[1]
    nk = lambda p: len(p["kernels"])  # noqa: E731
    scratch = lambda p: {f["name"] for f in p["fields"] if f["kind"] == "temp"}  # noqa: E731
    expect = {"fw_pgrad_f32": (4, 3, set()), "fuse_chain_f32": (6, 3, set()), "fuse_reuse_f64": (4, 2, {"t", "u"}),
              "fuse_partial_f64": (5, 3, {"t"}), "tmp_koffset_f64": (4, 4, {"t"}), "k_intervals_f64": (3, 3, set())}[name]  # fmt: skip
    assert (nk(plain), nk(fused), scratch(fused)) == expect
    for opts in ({}, {"fuse_loops": False}, {"interior_loop": True, "static_pitch": 160}):
        run_emulated(name, "staged", opts, domain=(139, 70, 4), seed=13, layout="b200", guard="end")
        run_emulated(name, "default", opts, domain=(21, 9, 3), seed=14)


def test_loop_fusion_legality_rules():
    """_can_fuse on hand-made loops: K-offset read of the producer's output, an offset read of a field
    the later loop overwrites, and interval bounds whose order depends on the domain size are refused"""
    from gt4py_b200 import codegen_stream as cs

    def fld(name, off=(0, 0, 0)):
        return {"t": "field", "name": name, "dtype": "float64", "off": list(off), "data_index": []}

    def loop(assigns, interval=(("start", 0), ("end", 0))):
        body = [{"t": "assign", "left": fld(left), "right": right} for left, right in assigns]
        return {"order": "parallel", "caches": [], "sections": [{"interval": [list(interval[0]), list(interval[1])],
                "hes": [{"locals": [], "extent": [[0, 0], [0, 0]], "body": body}]}]}  # fmt: skip

    st = {"domain_info": {"min_k": 2}}
    a = loop([("t", fld("x"))])
    assert cs._can_fuse(a, loop([("y", fld("t", (1, 0, 0)))]), st)  # IJ-offset hand-over: register windows
    assert not cs._can_fuse(a, loop([("y", fld("t", (0, 0, 1)))]), st)  # level k+1 of t may not exist yet
    assert not cs._can_fuse(loop([("t", fld("x", (0, 1, 0)))]), loop([("x", fld("t"))]), st)  # WAR on x
    assert cs._can_fuse(loop([("t", fld("x"))]), loop([("x", fld("t"))]), st)  # zero-offset read-modify-write is fine
    lo = loop([("t", fld("x"))], (("start", 0), ("start", 4)))
    hi = loop([("y", fld("t"))], (("end", -1), ("end", 0)))
    assert not cs._can_fuse(lo, hi, st)  # start+4 <= end-1 needs nK >= 5 > min_k
    assert cs._can_fuse(lo, hi, {"domain_info": {"min_k": 5}})
    fused = cs._fuse(lo, hi)
    assert [s["interval"] for s in fused["sections"]] == [[["start", 0], ["start", 4]], [["end", -1], ["end", 0]]]
    both = cs._fuse(loop([("t", fld("x"))]), loop([("y", fld("t"))], (("start", 1), ("end", -1))))
    assert [(s["interval"], len(s["hes"])) for s in both["sections"]] == [
        ([["start", 0], ["start", 1]], 1), ([["start", 1], ["end", -1]], 2), ([["end", -1], ["end", 0]], 1)]  # fmt: skip


# ---- hazards the reference's own checks never let through, but hand-built / fused IRs can ---------------------------------
def _inplace_diffusion_ir():
    """`t = x; x = t[-1,0,0] + t[1,0,0] + t[0,-1,0] + t[0,1,0]`: the incoming state of x is needed on the neighbours' cells
    while x is being overwritten.  gt4py rejects it at GTIR->OIR ("non-zero read extent on written fields",
    gtc/gtir_to_oir.py:19-46); an IR built by hand (or by a fusion tool) must still run correctly."""
    f64 = "float64"

    def fa(name, off=(0, 0, 0)):
        return {"t": "field", "name": name, "off": list(off), "data_index": [], "dtype": f64}

    def add(a, b):
        return {"t": "binary", "op": "+", "left": a, "right": b, "dtype": f64}

    hes = [
        {"body": [{"t": "assign", "left": fa("t"), "right": fa("x")}], "extent": [[-1, 1], [-1, 1]], "locals": []},
        {"body": [{"t": "assign", "left": fa("x"),
                   "right": add(add(add(fa("t", (-1, 0, 0)), fa("t", (1, 0, 0))), fa("t", (0, -1, 0))), fa("t", (0, 1, 0)))}],
         "extent": [[0, 0], [0, 0]], "locals": []},
    ]  # fmt: skip
    decl = {"data_dims": [], "dims": [True, True, True], "dtype": f64}
    return {
        "t": "stencil", "name": "inplace_diffuse_f64", "ir_version": 1, "variant": "staged", "options": {},
        "domain_info": {"min_k": 0}, "parameter_info": {},
        "field_info": {"x": {"access": "READ_WRITE", "axes": ["I", "J", "K"], "boundary": [[1, 1], [1, 1], [0, 0]], "data_dims": [], "dtype": f64}},
        "params": [{**decl, "name": "x", "t": "field"}],
        "temporaries": [{**decl, "name": "t", "extent": [[-1, 1], [-1, 1]]}],
        "loops": [{"order": "parallel", "caches": [], "sections": [{"interval": [["start", 0], ["end", 0]], "hes": hes}]}],
    }  # fmt: skip


@pytest.mark.parametrize("options", [{"strategy": "auto"}, {"strategy": "point"}, {"strategy": "auto", "interior_loop": True, "tma": 2}])
def test_in_place_update_through_a_temporary_on_a_multi_tile_domain(options):
    """the streaming generator kept `x` as an input stream with halo lanes / rows AND stored it in the
    same kernel: warps read cells their neighbours had already overwritten (1792 wrong cells on this domain).  Such a
    loop is no longer streamable; the point generator splits it at the hazard."""
    st = _inplace_diffusion_ir()
    domain = (300, 150, 2)
    rng = np.random.default_rng(3)
    x = rng.random((domain[0] + 2, domain[1] + 2, domain[2]))
    origins = {"x": (1, 1, 0)}
    ref = {"x": x.copy()}
    numpy_oracle.run(st, ref, {}, domain, origins)
    es = EmuStencil(st, options, name="inplace_diffuse")
    got = {"x": x.copy()}
    es.run(got, {}, domain, origins, layout="b200", guard="end")
    np.testing.assert_array_equal(got["x"], ref["x"])
    assert len(es.plan["kernels"]) >= 2 and all(k["kind"] != "stream" for k in es.plan["kernels"])


def test_parallel_sections_coupled_through_a_k_offset_are_separate_ordered_launches():
    """one kernel for all sections of a PARALLEL loop runs them concurrently (K on blockIdx.z);
    `interval(0,1): b = a*2; interval(1,None): c = b[0,0,-1] + 1` (one vertical loop after AdjacentLoopMerging) needs
    the first section finished before the second starts."""
    for variant in ("default", "staged"):
        st = testing.load_ir("sections_koff_f64", variant)
        assert len(st["loops"]) == 1 and len(st["loops"][0]["sections"]) == 2  # the reference merged the two computations
        for strategy in ("point", "auto"):
            es = run_emulated("sections_koff_f64", variant, {"strategy": strategy}, check_golden=True)
            launches = [s for s in es.plan["steps"] if s["t"] == "launch"]
            assert len(launches) == 2, (variant, strategy, es.plan["steps"])
    # independent sections still share one kernel
    es = run_emulated("k_intervals_f64", "default", {"strategy": "point"})
    assert len([s for s in es.plan["steps"] if s["t"] == "launch"]) == 1


def test_halo_wait_kernels_put_the_boundary_tiles_last_and_wait_for_the_flags():
    """multi-GPU variant of the streaming kernel (`halo_wait`): same results as the plain kernel; the tiles that read halo
    rows are the last tasks of the grid and wait for the neighbours' flags (Geom::halo_flag_*, b200_stencil_run_halo)"""
    import ctypes

    st = testing.load_ir("hdiff_f32", "staged")
    domain = (150, 260, 3)  # 5 J tiles of 64 rows
    fields, params, origins, domain = testing.make_case_data(st, "hdiff_f32", domain=domain, seed=4)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    flags = (ctypes.c_uint64 * 2)(7, 7)
    addr = ctypes.addressof(flags)
    for opts in ({"halo_wait": True}, {"halo_wait": True, "interior_loop": True, "static_pitch": 160},
                 {"halo_wait": True, "interior_loop": True, "static_pitch": 160, "tma": 2, "tile_j": 32, "tma_mode": "bulk"},
                 {"halo_wait": True, "interior_loop": True, "static_pitch": 160, "halo_lean": True}, {"halo_wait": True, "halo_lean": True}):
        es = EmuStencil(st, opts, name="hdiff_halo_wait")
        assert ("fence_proxy_async_global" in es.source) == bool(opts.get("tma"))  # flag acquire (generic proxy) -> bulk copies (async proxy)
        assert "wait_flag" in es.source and "interior tiles first" in es.source
        assert ("wait_flag_nolimit" in es.source) == bool(opts.get("halo_lean"))
        for hw in ((0, 0, 0), (addr, addr + 8, 7), (addr, 0, 5)):
            got = {k: v.copy() for k, v in fields.items()}
            es.run(got, params, domain, origins, layout="b200", guard="end", halo_wait=hw)
            np.testing.assert_array_equal(got["out_field"], ref["out_field"])
    assert "wait_flag" not in EmuStencil(st, {}, name="hdiff_plain").source


@pytest.mark.parametrize("opts", [{}, {"k_order": False}, {"halo_wait": True}, {"interior_loop": True, "static_pitch": 160, "halo_wait": True},
                                  {"tma": 2, "tile_j": 16, "interior_loop": True, "static_pitch": 160}])  # fmt: skip
def test_level_fastest_task_order_of_kernels_that_read_at_k_offsets(opts):
    """kernels with K-offset inputs take the LEVEL as the fastest task index (L1/L2 re-use of the planes read from three
    levels); one- and two-level sections use short J tiles.  Same results, every task decoded exactly once."""
    for name in ("fw_pgrad_f32", "fw_div_f32"):
        st = testing.load_ir(name, "staged")
        es = EmuStencil(st, {"strategy": "auto", **opts}, name=name + "_korder")
        ks = [k for k in es.plan["kernels"] if k["kind"] == "stream"]
        assert ks and (("task % nk" in es.source) == (opts.get("k_order", "auto") is not False))
        assert {k["tile"][1] for k in ks} == {8, opts.get("tile_j", 64)}  # thin top / bottom sections, full-height middle section
        assert not any(k["tma"] for k in ks if k["tile"][1] == 8)  # no shared-memory ring for the thin sections
        run_emulated(name, "staged", {"strategy": "auto", **opts}, domain=(139, 150, 5), seed=3, layout="b200", guard="end")


def test_fused_sweeps_keep_their_private_temporaries_in_shared_memory():
    """`col_smem`: the coefficients the forward sweep of the fast-waves w solver hands to its back substitution never
    touch global memory (one kernel, dynamic shared memory sized by the launcher per level); a domain taller than the
    shared-memory budget takes the global-scratch code of the same kernel."""
    st = testing.load_ir("fw_wsolve_f32", "default")
    es = EmuStencil(st, {"col_smem": True}, name="wsolve_smem")
    (k,) = es.plan["kernels"]
    assert k["smem_fields"] == ["ccol", "dcol"] and k["smem_per_k"] == 2 * 4 * 64 and k["smem_kcap"] == 112 and k["block"] == [32, 2, 1]
    body = es.source[es.source.index("if (A.g.nK <= 112)"):es.source.index("return;\n  }")]
    assert "s_ccol[" in body and "c_ccol[" not in body and "c_dcol[" not in body
    from gt4py_b200 import codegen

    assert f" {k['smem_per_k']} {k['smem_kcap']}\n" in codegen.plan_to_text(es.plan)  # what the launcher sizes the request from
    # API fields another kernel does not see are not candidates: the Thomas solver's sup / rhs are outputs
    assert not EmuStencil(testing.load_ir("tridiagonal_f64", "default"), {"col_smem": True}, name="tri_smem").plan["kernels"][0].get("smem_fields")
    for domain in ((21, 6, 30), (5, 3, 130)):  # shared-memory path / fallback path (130 levels > 112)
        run_emulated("fw_wsolve_f32", "default", {"col_smem": True}, domain=domain, seed=1, layout="b200", guard="end")


def test_divisions_by_launch_invariants_take_the_reciprocal_path():
    """upwind5 divides by 60*dx, 60*dy: the hoisted-reciprocal sequence must be what runs — also in the prologue rows and the
    out-of-domain lanes of a tile (window registers start at one: a zero dividend would send the group to the IEEE
    fallback) — and the fallback must still give the reference's bits when a field holds zeros / infinities."""
    es = run_emulated("upwind5_f32", "staged", {"interior_loop": True, "static_pitch": 160}, domain=(139, 70, 2), seed=5, layout="b200", guard="end")
    assert "b200::div_inv_try(" in es.source and "dv1 = b200::div_inv_make(" in es.source
    tr = es.trace()
    assert tr[3] > 0 and tr[6] == 0, tr
    st = testing.load_ir("upwind5_f32", "staged")
    fields, params, origins, domain = testing.make_case_data(st, "upwind5_f32", domain=(75, 41, 2), seed=3)
    fields["u"][5:40, 7:20, :] = 0.0
    fields["u"][50, 30, 1] = -0.0
    fields["v"][20:30, 3:9, 0] = np.inf
    fields["v"][60, 8, 1] = 1e-44  # subnormal
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    es = EmuStencil(st, {"strategy": "auto"}, name="upwind5_specials")
    got = {k: v.copy() for k, v in fields.items()}
    es.run(got, params, domain, origins, layout="b200", guard="end")
    assert es.trace()[6] > 0
    np.testing.assert_array_equal(got["out"].view(np.uint32), ref["out"].view(np.uint32))  # bits, signed zeros and NaNs included
