"""Cross-stencil fusion (gt4py_b200/fuse.py): a fused sequence must equal the sequence.

Reference semantics of a hand-over field that becomes a temporary = gt4py's own temporary semantics
(extent analysis, reference oir_optimizations/utils.py:250-330): the producer is computed wherever the
consumers read it.  So `fused(A, B)` on domain D is checked against the ORACLE running A on D grown
by B's read extent and then B on D — and the generated kernels (CPU emulator, storage layout between
guard pages) against the oracle of the fused IR."""

import numpy as np
import pytest

from gt4py_b200 import codegen, fuse, testing
from oracle import numpy_oracle

from emu.emu import EmuStencil


def _grow(origin, h):
    return tuple(o - h if a < 2 else o for a, o in enumerate(origin))


def _emulate(st, fields, params, domain, origins, expect, option_sets):
    for opts in option_sets:
        got = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
        EmuStencil(st, opts, name=st["name"]).run(got, params, domain, origins, layout="b200", guard="end")
        for name in testing.written_fields(st):
            np.testing.assert_array_equal(got[name], expect[name], err_msg=f"{st['name']} {opts}: {name}")


def test_two_hdiff_steps_in_one_pass_equal_two_calls():
    st = testing.load_ir("hdiff_f32", "staged")
    f2 = fuse.repeat(st, 2, carry=("in_field", "out_field"))
    assert f2["field_info"]["in_field"]["boundary"][:2] == [[4, 4], [4, 4]] and f2["field_info"]["coeff"]["boundary"][:2] == [[2, 2], [2, 2]]
    assert f2["field_info"]["in_field"]["access"] == "READ" and f2["field_info"]["out_field"]["access"] == "WRITE"
    assert [p["name"] for p in f2["params"]] == ["in_field", "coeff", "out_field"]
    fields, params, origins, domain = testing.make_case_data(f2, "hdiff_f32", domain=(70, 37, 3), seed=1)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(f2, ref, params, domain, origins)
    # the sequence: step 1 on the domain grown by step 2's read extent (2), step 2 on the domain
    mid = np.zeros_like(fields["in_field"])
    seq_out = fields["out_field"].copy()
    o1 = {n: _grow(origins["in_field" if n != "coeff" else "coeff"], 2) for n in ("in_field", "out_field", "coeff")}
    numpy_oracle.run(st, {"in_field": fields["in_field"].copy(), "out_field": mid, "coeff": fields["coeff"].copy()}, {},
                     (domain[0] + 4, domain[1] + 4, domain[2]), o1)  # fmt: skip
    o2 = {"in_field": origins["in_field"], "out_field": origins["out_field"], "coeff": origins["coeff"]}
    numpy_oracle.run(st, {"in_field": mid, "out_field": seq_out, "coeff": fields["coeff"].copy()}, {}, domain, o2)
    np.testing.assert_array_equal(ref["out_field"], seq_out)
    # one streaming kernel, every hand-over value and temporary in registers: 12 B/cell for two updates
    _src, plan = codegen.generate(f2, {"strategy": "auto"})
    assert [k["kind"] for k in plan["kernels"]] == ["stream"]
    assert all(f["kind"] in ("api", "dead") for f in plan["fields"])
    assert testing.algorithmic_bytes_per_cell(f2) == 12
    _emulate(f2, fields, params, domain, origins, ref,
             [{"strategy": "auto"}, {"strategy": "point"}, {"strategy": "auto", "interior_loop": True, "static_pitch": 128}, {"strategy": "auto", "vector_width": 4}])  # fmt: skip


def test_three_steps_and_ragged_domains():
    st = testing.load_ir("laplacian_f64", "default")
    names = [p["name"] for p in st["params"] if p["t"] == "field"]
    src = next(n for n in names if st["field_info"][n]["access"] == "READ")
    dst = next(n for n in names if st["field_info"][n]["access"] == "WRITE")
    f3 = fuse.repeat(st, 3, carry=(src, dst))
    assert f3["field_info"][src]["boundary"][:2] == [[3, 3], [3, 3]]
    for domain in ((1, 1, 1), (33, 5, 2), (130, 40, 2)):
        fields, params, origins, domain = testing.make_case_data(f3, "laplacian_f64", domain=domain, seed=3)
        ref = {k: v.copy() for k, v in fields.items()}
        numpy_oracle.run(f3, ref, params, domain, origins)
        cur = fields[src].copy()
        for n, grow in enumerate((2, 1, 0)):
            nxt = np.zeros_like(cur) if n < 2 else fields[dst].copy()
            org = {src: _grow(origins[src], grow), dst: _grow(origins[src] if n < 2 else origins[dst], grow)}
            numpy_oracle.run(st, {src: cur, dst: nxt}, {}, (domain[0] + 2 * grow, domain[1] + 2 * grow, domain[2]), org)
            cur = nxt
        np.testing.assert_array_equal(ref[dst], cur)
        _emulate(f3, fields, params, domain, origins, ref, [{"strategy": "auto"}, {"strategy": "point"}])


def test_different_stencils_with_scalars_hdiff_then_upwind():
    """out = upwind5(phi = hdiff(in, coeff), u, v; dt, dx, dy): hdiff is computed on the domain grown by 3"""
    a = testing.load_ir("hdiff_f32", "staged")
    b = testing.load_ir("upwind5_f32", "staged")
    fused = fuse.compose("hdiff_upwind", [(a, {"out_field": "phi"}), (b, {})], intermediates=["phi"])
    assert fused["field_info"]["in_field"]["boundary"][:2] == [[5, 5], [5, 5]]
    assert "phi" not in fused["field_info"] and set(fused["parameter_info"]) == {"dt", "dx", "dy"}
    fields, params, origins, domain = testing.make_case_data(fused, "upwind5_f32", domain=(45, 20, 2), seed=5)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(fused, ref, params, domain, origins)
    phi = np.zeros_like(fields["in_field"])
    numpy_oracle.run(a, {"in_field": fields["in_field"].copy(), "out_field": phi, "coeff": fields["coeff"].copy()}, {},
                     (domain[0] + 6, domain[1] + 6, domain[2]),
                     {"in_field": _grow(origins["in_field"], 3), "out_field": _grow(origins["in_field"], 3), "coeff": _grow(origins["coeff"], 3)})  # fmt: skip
    out = fields["out"].copy()
    numpy_oracle.run(b, {"phi": phi, "u": fields["u"].copy(), "v": fields["v"].copy(), "out": out}, params, domain,
                     {"phi": origins["in_field"], "u": origins["u"], "v": origins["v"], "out": origins["out"]})  # fmt: skip
    np.testing.assert_array_equal(ref["out"], out)
    _src, plan = codegen.generate(fused, {"strategy": "auto"})
    assert [k["kind"] for k in plan["kernels"]] == ["stream"] and all(f["kind"] in ("api", "dead") for f in plan["fields"])
    _emulate(fused, fields, params, domain, origins, ref, [{"strategy": "auto"}, {"strategy": "point"}])


def test_hand_over_field_kept_as_an_output_is_stored_on_the_grown_domain():
    """without `intermediates` the hand-over field stays an argument; it is computed AND stored wherever
    the consumer reads it — the reference's rule for a field written and re-read at an offset inside one
    stencil (compute_extents unions the extents of every written field, oir_optimizations/utils.py:293-313)"""
    st = testing.load_ir("laplacian_f64", "default")
    names = [p["name"] for p in st["params"] if p["t"] == "field"]
    src = next(n for n in names if st["field_info"][n]["access"] == "READ")
    dst = next(n for n in names if st["field_info"][n]["access"] == "WRITE")
    fused = fuse.compose("lap_lap_mem", [(st, {dst: "mid"}), (st, {src: "mid"})])
    assert fused["field_info"]["mid"]["access"] == "WRITE" and fused["field_info"]["mid"]["boundary"][:2] == [[1, 1], [1, 1]]
    assert fused["field_info"][src]["boundary"][:2] == [[2, 2], [2, 2]]
    fields, params, origins, domain = testing.make_case_data(fused, "laplacian_f64", domain=(40, 21, 2), seed=7)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(fused, ref, params, domain, origins)
    seq = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, {src: seq[src], dst: seq["mid"]}, {}, (domain[0] + 2, domain[1] + 2, domain[2]),
                     {src: _grow(origins[src], 1), dst: _grow(origins["mid"], 1)})  # fmt: skip
    numpy_oracle.run(st, {src: seq["mid"], dst: seq[dst]}, {}, domain, {src: origins["mid"], dst: origins[dst]})
    for n in ("mid", dst):
        np.testing.assert_array_equal(ref[n], seq[n])
    _emulate(fused, fields, params, domain, origins, ref, [{"strategy": "auto"}, {"strategy": "point"}])


def test_column_solver_followed_by_a_pointwise_stage():
    tri = testing.load_ir("tridiagonal_f64", "default")
    cp = testing.load_ir("copy_f64", "default")
    cnames = [p["name"] for p in cp["params"] if p["t"] == "field"]
    csrc = next(n for n in cnames if cp["field_info"][n]["access"] == "READ")
    cdst = next(n for n in cnames if cp["field_info"][n]["access"] == "WRITE")
    fused = fuse.compose("tri_copy", [(tri, {"out": "x"}), (cp, {csrc: "x", cdst: "result"})], intermediates=["x"])
    fields, params, origins, domain = testing.make_case_data(fused, "tridiagonal_f64", domain=(33, 9, 12), seed=9)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(fused, ref, params, domain, origins)
    seq = {k: v.copy() for k, v in fields.items()}
    x = np.zeros_like(seq["rhs"])
    numpy_oracle.run(tri, {**{n: seq[n] for n in ("inf", "diag", "sup", "rhs")}, "out": x}, {}, domain,
                     {**{n: origins[n] for n in ("inf", "diag", "sup", "rhs")}, "out": origins["rhs"]})  # fmt: skip
    np.testing.assert_array_equal(ref["result"][tuple(slice(o, o + d) for o, d in zip(origins["result"], domain))],
                                  x[tuple(slice(o, o + d) for o, d in zip(origins["rhs"], domain))])  # fmt: skip
    _emulate(fused, fields, params, domain, origins, ref, [{"strategy": "auto"}, {"strategy": "point"}])


def test_rejections():
    st = testing.load_ir("laplacian_f64", "default")
    names = [p["name"] for p in st["params"] if p["t"] == "field"]
    src = next(n for n in names if st["field_info"][n]["access"] == "READ")
    dst = next(n for n in names if st["field_info"][n]["access"] == "WRITE")
    with pytest.raises(ValueError, match="read before"):
        fuse.compose("bad", [(st, {src: "t", dst: "o"}), (st, {src: "o", dst: "t"})], intermediates=["t"])
    with pytest.raises(ValueError, match="no argument"):
        fuse.compose("bad", [(st, {"nope": "t"})])
    with pytest.raises(ValueError, match="not a field"):
        fuse.compose("bad", [(st, {})], intermediates=["zz"])
    regions = [c for c in testing.list_cases() if "region" in c]
    if regions:
        with pytest.raises(NotImplementedError, match="horizontal regions"):
            r = testing.load_ir(regions[0], "default")
            fuse.compose("bad", [(r, {}), (r, {})])
    f32 = testing.load_ir("hdiff_f32", "staged")
    with pytest.raises(TypeError, match="is used as"):
        fuse.compose("bad", [(st, {dst: "x"}), (f32, {"in_field": "x"})])


def _has_hregion(st):
    found = []

    def visit(stmts):
        for s in stmts:
            if s["t"] == "hregion":
                found.append(s)
            if "body" in s:
                visit(s["body"])

    for *_x, he in __import__("gt4py_b200.ir", fromlist=["iter_hes"]).iter_hes(st):
        visit(he["body"])
    return bool(found)


@pytest.mark.parametrize("variant", ["default", "staged"])
@pytest.mark.parametrize("name", testing.list_cases())
def test_extent_analysis_reproduces_the_reference_on_every_fixture(name, variant):
    """fuse.recompute_extents against the extents the REFERENCE's compute_extents produced for the golden
    IRs (tests/golden/ir, written by tools/make_golden.py from gt4py itself): block extents, temporaries'
    extents and the IJ boundary of every API field must come out identical."""
    import copy

    st = testing.load_ir(name, variant)
    if _has_hregion(st):
        pytest.skip("horizontal regions: position-dependent extents (not composed)")
    mine = copy.deepcopy(st)
    need = fuse.recompute_extents(mine)
    from gt4py_b200 import ir as b2ir

    for (*_a, he_ref), (*_b, he_mine) in zip(b2ir.iter_hes(st), b2ir.iter_hes(mine)):
        assert he_mine["extent"] == he_ref["extent"]
    used = {a["name"] for *_x, he in b2ir.iter_hes(st) for a in b2ir.field_accesses(he["body"])}
    for t_ref, t_mine in zip(st["temporaries"], mine["temporaries"]):
        if t_ref["name"] in used:
            assert t_mine["extent"] == t_ref["extent"], t_ref["name"]
    for fname, fi in st["field_info"].items():
        if fi is None or fi["access"] == "NONE":
            continue
        e = need[fname]
        got = [[-e[0][0], e[0][1]] if "I" in fi["axes"] else [0, 0], [-e[1][0], e[1][1]] if "J" in fi["axes"] else [0, 0]]
        # (a field accessed at one-sided offsets only has a NEGATIVE boundary on the other side in the reference, e.g. read
        # at [0,-1,0] alone -> J boundary (1, -1); the fused IR reports the boundary clamped at zero, which is sufficient)
        assert got == [[max(x, 0) for x in b] for b in fi["boundary"][:2]], fname


def _fusable_pairs(n, seed):
    """seeded sample of (producer fixture, consumer fixture, written field, read field) with matching dtypes"""
    import random

    pool = []
    for name in testing.list_cases():
        st = testing.load_ir(name, "staged")
        if _has_hregion(st) or "while" in name or name in ("math_f64", "math_f32", "rounding_f64") or "varoff" in name:
            continue
        pool.append((name, st))
    rng = random.Random(seed)
    pairs = [(a, b) for a in pool for b in pool]
    rng.shuffle(pairs)
    out = []
    for (na, a), (nb, b) in pairs:
        wa = [n for n, fi in a["field_info"].items() if fi and fi["access"] == "WRITE" and len(fi["axes"]) == 3 and not fi["data_dims"]]
        rb = [n for n, fi in b["field_info"].items()
              if fi and fi["access"] == "READ" and len(fi["axes"]) == 3 and not fi["data_dims"] and tuple(fi["boundary"][2]) == (0, 0)]  # fmt: skip
        ok = [(x, y) for x in wa for y in rb if a["field_info"][x]["dtype"] == b["field_info"][y]["dtype"]]
        if ok:
            out.append((na, nb, *rng.choice(ok)))
        if len(out) == n:
            break
    return out


@pytest.mark.parametrize("na,nb,x,y", _fusable_pairs(14, seed=2026))
def test_random_fixture_pairs_fused(na, nb, x, y):
    """differential check of the generators on fused IRs (seeded sample of what /tmp fuzzing ran 190 pairs of):
    producer.x -> consumer.y as an intermediate; emulated kernels (storage layout, guard pages) == oracle"""
    a, b = testing.load_ir(na, "staged"), testing.load_ir(nb, "staged")
    bind_b = {n: "b_" + n for n in [p["name"] for p in b["params"]] + list(b["field_info"]) + list(b["parameter_info"])}
    bind_b[y] = "handover"
    fused = fuse.compose(f"fz_{na}_{nb}", [(a, {x: "handover"}), (b, bind_b)], intermediates=["handover"])
    domain = (37, 9, max(4, int(fused["domain_info"]["min_k"])))
    fields, params, origins, domain = testing.make_case_data(fused, fused["name"], domain=domain, seed=11)  # no fixture recipe
    ref = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
    with np.errstate(all="ignore"):
        numpy_oracle.run(fused, ref, params, domain, origins)
    for opts in ({"strategy": "auto"}, {"strategy": "point"}):
        got = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
        EmuStencil(fused, opts, name=fused["name"]).run(got, params, domain, origins, layout="b200", guard="end")
        for n in testing.written_fields(fused):
            np.testing.assert_array_equal(got[n], ref[n], err_msg=f"{na}->{nb} {opts}: {n}")
