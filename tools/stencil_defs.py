"""GTScript definitions used for fixtures, parity tests and benchmarks (needs the gt4py frontend).

DEV TOOL: imported only by `tools/make_golden.py` (in the build container, where the reference is
mounted) and by the optional full-API tests.  The lowered IR of every entry is committed under
`tests/golden/ir/` so the GPU box never needs this file.

The benchmark stencils restate the algorithms the reference names in BASELINE.json:
  * hdiff                  tests/cartesian_tests/integration_tests/multi_feature_tests/stencil_definitions.py:316-328
  * hdiff_simple           .../stencil_definitions.py:206-216
  * laplacian              .../test_suites.py:233-236  (lap_op)
  * tridiagonal            .../stencil_definitions.py:219-232
  * vadv (dycore)          .../stencil_definitions.py:235-313
  * upwind5 advection      examples/cartesian/demo_burgers.ipynb cell 6 (advection_x / advection_y)
The remaining entries exercise one OIR feature each.
"""


import numpy as np

from gt4py.cartesian import gtscript
from gt4py.cartesian.gtscript import (  # noqa: F401
    BACKWARD,
    FORWARD,
    IJ,
    IJK,
    K,
    PARALLEL,
    Field,
    I,
    J,
    computation,
    horizontal,
    interval,
    region,
)

F32 = Field[np.float32]
F64 = Field[np.float64]
I32 = Field[np.int32]
I64 = Field[np.int64]
B8 = Field[np.bool_]

REGISTRY = {}


def case(name=None, *, build=None, externals=None, variants=("default", "staged")):
    def deco(fn):
        REGISTRY[name or fn.__name__] = {
            "definition": fn,
            "build": dict(build or {}),
            "externals": externals,
            "variants": tuple(variants),
        }
        return fn

    return deco


# ---------------------------------------------------------------------------------------------
# Benchmark stencils
# ---------------------------------------------------------------------------------------------
def _hdiff_body():
    pass


@case("hdiff_f32", build={"literal_float_precision": 32})
def hdiff_f32(in_field: F32, out_field: F32, coeff: F32):
    with computation(PARALLEL), interval(...):
        lap_field = 4.0 * in_field[0, 0, 0] - (
            in_field[1, 0, 0] + in_field[-1, 0, 0] + in_field[0, 1, 0] + in_field[0, -1, 0]
        )
        res = lap_field[1, 0, 0] - lap_field[0, 0, 0]
        flx_field = 0 if (res * (in_field[1, 0, 0] - in_field[0, 0, 0])) > 0 else res
        res = lap_field[0, 1, 0] - lap_field[0, 0, 0]
        fly_field = 0 if (res * (in_field[0, 1, 0] - in_field[0, 0, 0])) > 0 else res
        out_field = in_field[0, 0, 0] - coeff[0, 0, 0] * (
            flx_field[0, 0, 0] - flx_field[-1, 0, 0] + fly_field[0, 0, 0] - fly_field[0, -1, 0]
        )


@case("hdiff_f32_default_literals")
def hdiff_f32_default_literals(in_field: F32, out_field: F32, coeff: F32):
    # same stencil, default literal precision: every temporary is FLOAT64
    with computation(PARALLEL), interval(...):
        lap_field = 4.0 * in_field[0, 0, 0] - (
            in_field[1, 0, 0] + in_field[-1, 0, 0] + in_field[0, 1, 0] + in_field[0, -1, 0]
        )
        res = lap_field[1, 0, 0] - lap_field[0, 0, 0]
        flx_field = 0 if (res * (in_field[1, 0, 0] - in_field[0, 0, 0])) > 0 else res
        res = lap_field[0, 1, 0] - lap_field[0, 0, 0]
        fly_field = 0 if (res * (in_field[0, 1, 0] - in_field[0, 0, 0])) > 0 else res
        out_field = in_field[0, 0, 0] - coeff[0, 0, 0] * (
            flx_field[0, 0, 0] - flx_field[-1, 0, 0] + fly_field[0, 0, 0] - fly_field[0, -1, 0]
        )


@case("hdiff_f64")
def hdiff_f64(in_field: F64, out_field: F64, coeff: F64):
    with computation(PARALLEL), interval(...):
        lap_field = 4.0 * in_field[0, 0, 0] - (
            in_field[1, 0, 0] + in_field[-1, 0, 0] + in_field[0, 1, 0] + in_field[0, -1, 0]
        )
        res = lap_field[1, 0, 0] - lap_field[0, 0, 0]
        flx_field = 0 if (res * (in_field[1, 0, 0] - in_field[0, 0, 0])) > 0 else res
        res = lap_field[0, 1, 0] - lap_field[0, 0, 0]
        fly_field = 0 if (res * (in_field[0, 1, 0] - in_field[0, 0, 0])) > 0 else res
        out_field = in_field[0, 0, 0] - coeff[0, 0, 0] * (
            flx_field[0, 0, 0] - flx_field[-1, 0, 0] + fly_field[0, 0, 0] - fly_field[0, -1, 0]
        )


@case("hdiff_simple_f64")
def hdiff_simple_f64(in_field: F64, coeff: F64, out_field: F64):
    with computation(PARALLEL), interval(...):
        lap_field = 4.0 * in_field[0, 0, 0] - (
            in_field[1, 0, 0] + in_field[-1, 0, 0] + in_field[0, 1, 0] + in_field[0, -1, 0]
        )
        flx_field = lap_field[1, 0, 0] - lap_field[0, 0, 0]
        fly_field = lap_field[0, 1, 0] - lap_field[0, 0, 0]
        out_field = in_field[0, 0, 0] - coeff[0, 0, 0] * (
            flx_field[0, 0, 0] - flx_field[-1, 0, 0] + fly_field[0, 0, 0] - fly_field[0, -1, 0]
        )


@case("laplacian_f64")
def laplacian_f64(u: F64, out: F64):
    with computation(PARALLEL), interval(...):
        out = 4.0 * u[0, 0, 0] - (u[1, 0, 0] + u[-1, 0, 0] + u[0, 1, 0] + u[0, -1, 0])


@case("tridiagonal_f64")
def tridiagonal_f64(inf: F64, diag: F64, sup: F64, rhs: F64, out: F64):
    with computation(FORWARD):
        with interval(0, 1):
            sup = sup / diag
            rhs = rhs / diag
        with interval(1, None):
            sup = sup / (diag - sup[0, 0, -1] * inf)
            rhs = (rhs - inf * rhs[0, 0, -1]) / (diag - sup[0, 0, -1] * inf)
    with computation(BACKWARD):
        with interval(-1, None):
            out = rhs
        with interval(0, -1):
            out = rhs - sup * out[0, 0, 1]


@case("vadv_f64", externals={"BET_M": 0.5, "BET_P": 0.5})
def vadv_f64(
    utens_stage: F64,
    u_stage: F64,
    wcon: F64,
    u_pos: F64,
    utens: F64,
    *,
    dtr_stage: float,
):
    from __externals__ import BET_M, BET_P

    with computation(FORWARD):
        with interval(0, 1):
            gcv = 0.25 * (wcon[1, 0, 1] + wcon[0, 0, 1])
            cs = gcv * BET_M
            ccol = gcv * BET_P
            bcol = dtr_stage - ccol[0, 0, 0]
            correction_term = -cs * (u_stage[0, 0, 1] - u_stage[0, 0, 0])
            dcol = (
                dtr_stage * u_pos[0, 0, 0] + utens[0, 0, 0] + utens_stage[0, 0, 0] + correction_term
            )
            divided = 1.0 / bcol[0, 0, 0]
            ccol = ccol[0, 0, 0] * divided
            dcol = dcol[0, 0, 0] * divided
        with interval(1, -1):
            gav = -0.25 * (wcon[1, 0, 0] + wcon[0, 0, 0])
            gcv = 0.25 * (wcon[1, 0, 1] + wcon[0, 0, 1])
            as_ = gav * BET_M
            cs = gcv * BET_M
            acol = gav * BET_P
            ccol = gcv * BET_P
            bcol = dtr_stage - acol[0, 0, 0] - ccol[0, 0, 0]
            correction_term = -as_ * (u_stage[0, 0, -1] - u_stage[0, 0, 0]) - cs * (
                u_stage[0, 0, 1] - u_stage[0, 0, 0]
            )
            dcol = (
                dtr_stage * u_pos[0, 0, 0] + utens[0, 0, 0] + utens_stage[0, 0, 0] + correction_term
            )
            divided = 1.0 / (bcol[0, 0, 0] - ccol[0, 0, -1] * acol[0, 0, 0])
            ccol = ccol[0, 0, 0] * divided
            dcol = (dcol[0, 0, 0] - (dcol[0, 0, -1]) * acol[0, 0, 0]) * divided
        with interval(-1, None):
            gav = -0.25 * (wcon[1, 0, 0] + wcon[0, 0, 0])
            as_ = gav * BET_M
            acol = gav * BET_P
            bcol = dtr_stage - acol[0, 0, 0]
            correction_term = -as_ * (u_stage[0, 0, -1] - u_stage[0, 0, 0])
            dcol = (
                dtr_stage * u_pos[0, 0, 0] + utens[0, 0, 0] + utens_stage[0, 0, 0] + correction_term
            )
            divided = 1.0 / (bcol[0, 0, 0] - ccol[0, 0, -1] * acol[0, 0, 0])
            dcol = (dcol[0, 0, 0] - (dcol[0, 0, -1]) * acol[0, 0, 0]) * divided
    with computation(BACKWARD):
        with interval(-1, None):
            datacol = dcol[0, 0, 0]
            utens_stage = dtr_stage * (datacol - u_pos[0, 0, 0])
        with interval(0, -1):
            datacol = dcol[0, 0, 0] - ccol[0, 0, 0] * datacol[0, 0, 1]
            utens_stage = dtr_stage * (datacol - u_pos[0, 0, 0])


@gtscript.function
def _absval(phi):
    return phi[0, 0, 0] * (phi[0, 0, 0] >= 0.0) - phi[0, 0, 0] * (phi[0, 0, 0] < 0.0)


@gtscript.function
def _upwind5_x(dx, u, abs_u, phi):
    return u[0, 0, 0] / (60.0 * dx) * (
        +45.0 * (phi[1, 0, 0] - phi[-1, 0, 0])
        - 9.0 * (phi[2, 0, 0] - phi[-2, 0, 0])
        + (phi[3, 0, 0] - phi[-3, 0, 0])
    ) - abs_u[0, 0, 0] / (60.0 * dx) * (
        +(phi[3, 0, 0] + phi[-3, 0, 0])
        - 6.0 * (phi[2, 0, 0] + phi[-2, 0, 0])
        + 15.0 * (phi[1, 0, 0] + phi[-1, 0, 0])
        - 20.0 * phi[0, 0, 0]
    )


@gtscript.function
def _upwind5_y(dy, v, abs_v, phi):
    return v[0, 0, 0] / (60.0 * dy) * (
        +45.0 * (phi[0, 1, 0] - phi[0, -1, 0])
        - 9.0 * (phi[0, 2, 0] - phi[0, -2, 0])
        + (phi[0, 3, 0] - phi[0, -3, 0])
    ) - abs_v[0, 0, 0] / (60.0 * dy) * (
        +(phi[0, 3, 0] + phi[0, -3, 0])
        - 6.0 * (phi[0, 2, 0] + phi[0, -2, 0])
        + 15.0 * (phi[0, 1, 0] + phi[0, -1, 0])
        - 20.0 * phi[0, 0, 0]
    )


@case("upwind5_f32", build={"literal_float_precision": 32})
def upwind5_f32(
    phi: F32, u: F32, v: F32, out: F32, *, dt: np.float32, dx: np.float32, dy: np.float32
):
    with computation(PARALLEL), interval(...):
        abs_u = _absval(u)
        abs_v = _absval(v)
        adv_x = _upwind5_x(dx, u, abs_u, phi)
        adv_y = _upwind5_y(dy, v, abs_v, phi)
        out = phi[0, 0, 0] - dt * (adv_x[0, 0, 0] + adv_y[0, 0, 0])


# ---------------------------------------------------------------------------------------------
# Config 5: COSMO-style fast-waves / pressure-gradient suite (authored here; not in the reference)
# ---------------------------------------------------------------------------------------------
@case("fw_pgrad_f32", build={"literal_float_precision": 32})
def fw_pgrad_f32(
    pp: F32, rho: F32, hhl: F32, u_in: F32, v_in: F32, u_out: F32, v_out: F32, *, dt: np.float32, edadlat: np.float32
):
    """Horizontal pressure-gradient update of u, v (terrain-following correction with K±1)."""
    with computation(PARALLEL):
        with interval(0, 1):
            dpdz = (pp[0, 0, 1] - pp[0, 0, 0]) / (hhl[0, 0, 1] - hhl[0, 0, 0])
        with interval(1, -1):
            dpdz = (pp[0, 0, 1] - pp[0, 0, -1]) / (hhl[0, 0, 1] - hhl[0, 0, -1])
        with interval(-1, None):
            dpdz = (pp[0, 0, 0] - pp[0, 0, -1]) / (hhl[0, 0, 0] - hhl[0, 0, -1])
    with computation(PARALLEL), interval(...):
        dzdx = 0.5 * (hhl[1, 0, 0] - hhl[0, 0, 0])
        dzdy = 0.5 * (hhl[0, 1, 0] - hhl[0, 0, 0])
        pgx = (pp[1, 0, 0] - pp[0, 0, 0]) - dzdx * (dpdz[1, 0, 0] + dpdz[0, 0, 0])
        pgy = (pp[0, 1, 0] - pp[0, 0, 0]) - dzdy * (dpdz[0, 1, 0] + dpdz[0, 0, 0])
        u_out = u_in[0, 0, 0] - dt * edadlat * 2.0 * pgx / (rho[1, 0, 0] + rho[0, 0, 0])
        v_out = v_in[0, 0, 0] - dt * edadlat * 2.0 * pgy / (rho[0, 1, 0] + rho[0, 0, 0])


@case("fw_div_f32", build={"literal_float_precision": 32})
def fw_div_f32(u: F32, v: F32, w: F32, hhl: F32, div: F32, *, edadlat: np.float32):
    """3-D divergence used by the fast-waves pressure update."""
    with computation(PARALLEL):
        with interval(0, -1):
            dz = hhl[0, 0, 1] - hhl[0, 0, 0]
            div = edadlat * ((u[0, 0, 0] - u[-1, 0, 0]) + (v[0, 0, 0] - v[0, -1, 0])) + (
                w[0, 0, 1] - w[0, 0, 0]
            ) / dz
        with interval(-1, None):
            div = edadlat * ((u[0, 0, 0] - u[-1, 0, 0]) + (v[0, 0, 0] - v[0, -1, 0]))


@case("fw_wsolve_f32", build={"literal_float_precision": 32})
def fw_wsolve_f32(pp: F32, div: F32, rho: F32, w: F32, pp_out: F32, *, dt: np.float32, c2: np.float32):
    """Vertically implicit w / pp' update: tridiagonal sweep per column (FORWARD then BACKWARD)."""
    with computation(FORWARD):
        with interval(0, 1):
            a = 0.0
            b = 1.0 + dt * c2
            cc = -dt * c2 * 0.5
            ccol = cc / b
            dcol = (w[0, 0, 0] - dt * (pp[0, 0, 1] - pp[0, 0, 0]) / rho[0, 0, 0]) / b
        with interval(1, -1):
            a = -dt * c2 * 0.5
            b = 1.0 + dt * c2
            cc = -dt * c2 * 0.5
            den = b - a * ccol[0, 0, -1]
            ccol = cc / den
            dcol = ((w[0, 0, 0] - dt * (pp[0, 0, 1] - pp[0, 0, -1]) / (2.0 * rho[0, 0, 0])) - a * dcol[0, 0, -1]) / den
        with interval(-1, None):
            a = -dt * c2 * 0.5
            b = 1.0 + dt * c2
            den = b - a * ccol[0, 0, -1]
            ccol = 0.0
            dcol = ((w[0, 0, 0] - dt * (pp[0, 0, 0] - pp[0, 0, -1]) / rho[0, 0, 0]) - a * dcol[0, 0, -1]) / den
    with computation(BACKWARD):
        with interval(-1, None):
            w = dcol
            pp_out = pp[0, 0, 0] - dt * c2 * rho[0, 0, 0] * div[0, 0, 0]
        with interval(0, -1):
            w = dcol[0, 0, 0] - ccol[0, 0, 0] * w[0, 0, 1]
            pp_out = pp[0, 0, 0] - dt * c2 * rho[0, 0, 0] * (div[0, 0, 0] + (w[0, 0, 1] - w[0, 0, 0]))


# ---------------------------------------------------------------------------------------------
# Feature stencils
# ---------------------------------------------------------------------------------------------
@case("copy_f64")
def copy_f64(a: F64, b: F64):
    with computation(PARALLEL), interval(...):
        b = a


@case("scale_param_f32", build={"literal_float_precision": 32})
def scale_param_f32(a: F32, b: F32, *, alpha: np.float32, n: np.int32):
    with computation(PARALLEL), interval(...):
        b = alpha * a + n


@case("div_param_f32", build={"literal_float_precision": 32})
def div_param_f32(a: F32, b: F32, out: F32, out2: F32, *, dx: np.float32, dy: np.float32):
    """Divisions by launch invariants (scalar parameters, literals): the b200 generators divide through a hoisted
    reciprocal; the inputs of this case hold zeros of both signs, infinities, NaNs, subnormals and huge values."""
    with computation(PARALLEL), interval(...):
        out = a[1, 0, 0] / (60.0 * dx) - b[0, -1, 0] / dy + a[0, 0, 0] / 3.0
        out2 = (a[0, 0, 0] - b[0, 0, 0]) / (dx * dy) / 7.0


@case("div_param_col_f64")
def div_param_col_f64(a: F64, out: F64, *, dz: np.float64):
    with computation(FORWARD):
        with interval(0, 1):
            out = a / dz
        with interval(1, None):
            out = out[0, 0, -1] / 3.0 + a / (2.0 * dz)


@case("k_intervals_f64")
def k_intervals_f64(a: F64, b: F64):
    with computation(PARALLEL):
        with interval(0, 2):
            b = a
        with interval(2, -3):
            b = a + 1
        with interval(-3, None):
            b = a - 1


@case("if_field_f64")
def if_field_f64(a: F64, b: F64, c: F64):
    with computation(PARALLEL), interval(...):
        if a > 0.5:
            b = a * 2.0
            c = 1.0
        elif a > 0.25:
            b = -a
            c = 2.0
        else:
            b = 0.0
            c = c + a


@case("if_scalar_f64")
def if_scalar_f64(a: F64, b: F64, *, flag: np.int32):
    with computation(PARALLEL), interval(...):
        if flag > 0:
            b = a[1, 0, 0] + a[-1, 0, 0]
        else:
            b = a[0, 1, 0] - a[0, -1, 0]


@case("while_f64")
def while_f64(a: F64, b: F64):
    with computation(PARALLEL), interval(...):
        x = a
        n = 0.0
        while x < 1.0:
            n = n + 1.0
            x = x * 1.5 + 0.01
        b = n + x


@case("while_first_f64")
def while_first_f64(a: F64, b: F64):
    # the loop variable is updated FIRST: the reference numpy backend re-evaluates the loop condition as the mask of
    # every body statement (gtc/numpy/oir_to_npir.py:176-185), so `n` misses the increment of the iteration in which
    # x crosses 1.0 — unlike a per-point `while`.  north_star's oracle is the numpy backend: b200 reproduces it.
    with computation(PARALLEL), interval(...):
        x = a
        n = 0.0
        while x < 1.0:
            x = x * 1.5 + 0.01
            n = n + 1.0
        b = n + x


@case("while_masked_f64")
def while_masked_f64(a: F64, b: F64):
    with computation(PARALLEL), interval(...):
        x = a
        n = 0.0
        if a > 0.25:
            while x < 2.0:
                x = x * 1.25 + 0.125
                if x > 1.0:
                    n = n + 2.0
                else:
                    n = n + 1.0
        b = n * 10.0 + x


@case("regions_f64")
def regions_f64(a: F64, b: F64):
    with computation(PARALLEL), interval(...):
        b = a
        with horizontal(region[I[0], :], region[I[-1], :]):
            b = a * 10.0
        with horizontal(region[:, J[0] : J[0] + 2]):
            b = b + 100.0
        with horizontal(region[I[0] + 1 : I[-1], J[-1]]):
            b = -a


@case("region_extend_f64")
def region_extend_f64(a: F64, b: F64):
    with computation(PARALLEL), interval(...):
        t = a * 2.0
        with horizontal(region[I[0] - 1, :]):
            t = a * 3.0
        b = t[-1, 0, 0] + t[1, 0, 0]


@case("varoff_f64")
def varoff_f64(a: F64, idx: I32, b: F64):
    with computation(PARALLEL), interval(...):
        b = a[0, 0, idx] + a[0, 0, 1 - idx]


@case("lowdim_f64")
def lowdim_f64(a: F64, sfc: Field[IJ, np.float64], prof: Field[K, np.float64], b: F64):
    with computation(PARALLEL), interval(...):
        b = a * sfc + prof[0] + sfc[1, 0] * prof[1]


@case("lowdim_write_f64")
def lowdim_write_f64(a: F64, sfc: Field[IJ, np.float64]):
    with computation(FORWARD):
        with interval(0, 1):
            sfc = a
        with interval(1, None):
            sfc = sfc + a


@case("datadims_f64")
def datadims_f64(vec: Field[IJK, (np.float64, (3,))], mat: Field[IJK, (np.float64, (2, 2))], out: F64):
    with computation(PARALLEL), interval(...):
        out = vec[0, 0, 0][0] * mat[0, 0, 0][0, 0] + vec[1, 0, 0][1] * mat[0, 0, 0][0, 1] + vec[0, 0, 0][2] * mat[0, -1, 0][1, 1]


@case("datadims_write_f64")
def datadims_write_f64(a: F64, vec: Field[IJK, (np.float64, (2,))]):
    with computation(PARALLEL), interval(...):
        vec[0, 0, 0][0] = a
        vec[0, 0, 0][1] = -a + vec[0, 0, 0][0]


@case("ints_bools")
def ints_bools(a: I32, b: I64, m: B8, out_i: I64, out_m: B8):
    with computation(PARALLEL), interval(...):
        out_m = (a > 3) and (not m)
        out_i = a * 7 + b - (a - 2) * 3
        if m or (b < 0):
            out_i = -out_i


@case("kiter_f64")
def kiter_f64(a: F64, b: F64):
    with computation(PARALLEL), interval(...):
        b = a * K


@case("math_f64")
def math_f64(a: F64, o1: F64, o2: F64, o3: F64, o4: F64):
    with computation(PARALLEL), interval(...):
        o1 = sin(a) + cos(a) * tan(a * 0.5) + exp(-a) + log(a + 1.5) + sqrt(abs(a))
        o2 = floor(a * 3.7) + ceil(a * 2.2) + trunc(a * 5.0) + (a * 11.0) % 3.0 + (-a * 11.0) % 3.0
        o3 = min(a, 0.3) + max(a, 0.7) + a**2 + a**0.5 + abs(a - 0.5) ** 3.0
        o4 = asin(a * 0.9) + acos(a * 0.9) + atan(a) + asinh(a) + acosh(a + 1.0) + atanh(a * 0.9) + sinh(a) + cosh(a) + tanh(a) + cbrt(a) + log10(a + 1.0)


@case("math_f32", build={"literal_float_precision": 32})
def math_f32(a: F32, o1: F32, o2: F32):
    with computation(PARALLEL), interval(...):
        o1 = sqrt(abs(a)) + a / (a + 1.5) + abs(a - 0.5)
        o2 = floor(a * 3.7) + ceil(a * 2.2) + trunc(a * 5.0) + min(a, 0.3) + max(a, 0.7)


@case("rounding_f64")
def rounding_f64(a: F64, o1: F64, o2: F64, o3: F64):
    with computation(PARALLEL), interval(...):
        o1 = round(a * 10.0 - 5.0)
        o2 = round_away_from_zero(a * 10.0 - 5.0)
        o3 = erf(a) + erfc(a) + gamma(a + 0.5)


@case("isfinite_f64")
def isfinite_f64(a: F64, o: F64):
    with computation(PARALLEL), interval(...):
        x = 1.0 / (a - a)
        o = 1.0 if isinf(x) else (2.0 if isnan(x) else 3.0)
        if isfinite(a):
            o = o + 10.0


@case("casts")
def casts(a: F64, iv: I32, o_f: F32, o_i: I64, o_d: F64):
    with computation(PARALLEL), interval(...):
        o_f = a + iv
        o_i = int64(a * 100.0) + iv
        o_d = float32(a) * 3.0 + int32(a * 7.9)


@case("tmp_koffset_f64")
def tmp_koffset_f64(a: F64, b: F64):
    with computation(PARALLEL), interval(...):
        t = a * 2.0
    with computation(PARALLEL):
        with interval(0, 1):
            b = t[0, 0, 1]
        with interval(1, -1):
            b = t[0, 0, 1] - t[0, 0, -1]
        with interval(-1, None):
            b = -t[0, 0, -1]


@case("fwd_scan_f64")
def fwd_scan_f64(a: F64, acc: F64):
    with computation(FORWARD):
        with interval(0, 1):
            acc = a
        with interval(1, None):
            acc = acc[0, 0, -1] + a
    with computation(BACKWARD):
        with interval(-1, None):
            acc = acc * 2.0
        with interval(0, -1):
            acc = acc + acc[0, 0, 1] * 0.5


@case("fwd_tmp_ij_f64")
def fwd_tmp_ij_f64(a: F64, b: F64):
    # K-sequential loop whose temporary is read at an IJ offset in the same level
    with computation(FORWARD), interval(...):
        t = a * 2.0
        b = t[1, 0, 0] + t[-1, 0, 0] + t[0, 1, 0]


@case("two_stage_par_f32", build={"literal_float_precision": 32})
def two_stage_par_f32(a: F32, b: F32, c: F32):
    with computation(PARALLEL), interval(...):
        t = a[1, 0, 0] - a[-1, 0, 0]
        s = a[0, 1, 0] - a[0, -1, 0]
        b = t[0, 1, 0] + t[0, -1, 0] + s[1, 0, 0] * s[-1, 0, 0]
        c = b + t


@case("stage_halo_f32", build={"literal_float_precision": 32})
def stage_halo_f32(a: F32, b: F32, o1: F32, o2: F32):
    # an intermediate stage read at +1 in I while no INPUT is read at a positive I offset: the streaming kernel needs a
    # compute-only halo lane on the right (regression of a tools/fuzz_codegen.py finding: halo lanes were derived from
    # the inputs' reach only, so the last owned column of every warp segment took its own value for the neighbour's)
    with computation(PARALLEL), interval(...):
        t0 = a[0, 0, 0]
        t1 = t0[-1, -1, 0]
    with computation(PARALLEL), interval(...):
        o1 = t0[-2, 0, 0] / (abs(a[0, 0, 0]) + 1.0)
        o2 = b[0, -1, 0] if (t1[1, 2, 0] < t0[0, 0, 0]) else t1[0, 1, 0]


# ---------------------------------------------------------------------------------------------
# Column (FORWARD/BACKWARD) data-flow cases for the register k-cache generator (codegen_column.py)
# ---------------------------------------------------------------------------------------------
@case("col_mask_f64")
def col_mask_f64(a: F64, b: F64, c: F64):
    # masked writes / reads inside branches of fields that are carried along K
    with computation(FORWARD):
        with interval(0, 1):
            b = a
            c = 0.25
        with interval(1, None):
            if a > 0.5:
                b = b[0, 0, -1] + a
                c = a[0, 0, -1]
            else:
                b = a - c[0, 0, -1]
            c = c[0, 0, -1] * 0.5 + b
            if b > 1.0:
                c = c - b[0, 0, -1]


@case("col_chain_f64")
def col_chain_f64(a: F64, b: F64):
    # offsets -2 / +1 / +2 (pass-through registers), an IJ-offset read of a read-only field
    with computation(FORWARD):
        with interval(0, 2):
            b = a
        with interval(2, -2):
            b = b[0, 0, -2] + a[0, 0, -2] + a[0, 0, 2] * a[0, 0, 1] + a[1, 0, -1]
        with interval(-2, None):
            b = b[0, 0, -1] - b[0, 0, -2]


@case("col_backward_f64")
def col_backward_f64(a: F64, idx: I32, b: F64, c: F64):
    # BACKWARD sweep: carried b[0,0,1], a horizontal region, a variable-K read of the swept field
    with computation(BACKWARD):
        with interval(-1, None):
            b = a
            c = a
        with interval(0, -1):
            b = a + b[0, 0, 1] * 0.5
            with horizontal(region[I[0], :], region[:, J[-1]]):
                b = b * 2.0
            c = b[0, 0, idx] + b
            b = b + 1.0


@case("col_multiwrite_f32", build={"literal_float_precision": 32})
def col_multiwrite_f32(a: F32, b: F32, *, w: np.float32):
    # several writes of the same cell per level (dead stores), a temporary carried along K
    with computation(FORWARD):
        with interval(0, 1):
            t = a * w
            b = t
        with interval(1, None):
            t = t[0, 0, -1] * w + a
            b = t
            b = b * b - a
            if t > 1.0:
                b = b + t[0, 0, -1]
            t = t - 0.125


# ---------------------------------------------------------------------------------------------
# Several PARALLEL computations in one stencil: loop fusion by interval refinement (codegen_stream.py)
# ---------------------------------------------------------------------------------------------
@case("fuse_chain_f32", build={"literal_float_precision": 32})
def fuse_chain_f32(a: F32, b: F32, c: F32):
    # three computations with different vertical intervals; t and b are handed on at IJ offsets
    with computation(PARALLEL), interval(...):
        t = a[1, 0, 0] - a[0, 0, 0]
    with computation(PARALLEL):
        with interval(0, 1):
            s = t[0, 0, 0] * 2.0
        with interval(1, None):
            s = t[0, 1, 0] + a[0, 0, -1]
    with computation(PARALLEL):
        with interval(0, -1):
            c = s[0, 0, 0] + t[-1, 0, 0] + s[0, -1, 0]
        with interval(-1, None):
            c = s[1, 0, 0] - t[0, 0, 0]
    with computation(PARALLEL), interval(...):
        b = c[0, 0, 0] * 0.5 + s[0, 0, 0]


@case("fuse_reuse_f64")
def fuse_reuse_f64(a: F64, b: F64):
    # a temporary that is redefined by a later computation after being read at an offset, and a
    # masked redefinition in between
    with computation(PARALLEL), interval(...):
        t = a * 3.0
    with computation(PARALLEL), interval(...):
        u = t[1, 0, 0] + t[0, -1, 0]
    with computation(PARALLEL), interval(...):
        if a > 0.5:
            t = u * 2.0
    with computation(PARALLEL), interval(...):
        b = t[0, 0, 0] + u[0, 1, 0] - t[-1, 0, 0]


@case("fuse_partial_f64")
def fuse_partial_f64(a: F64, b: F64, c: F64):
    # the second computation reads the first one's result at a K offset (no fusion there), the third
    # and fourth can be fused with each other
    with computation(PARALLEL), interval(...):
        t = a + 1.0
    with computation(PARALLEL):
        with interval(0, 1):
            u = t[0, 0, 1]
        with interval(1, None):
            u = t[0, 0, -1] + t[0, 0, 0]
    with computation(PARALLEL), interval(...):
        v = u[1, 0, 0] * t[0, 1, 0]
    with computation(PARALLEL), interval(...):
        b = v[0, -1, 0] + u[0, 0, 0]
        c = v[-1, 0, 0]


@case("sections_koff_f64")
def sections_koff_f64(a: F64, b: F64, c: F64):
    # adjacent-interval PARALLEL computations are merged into ONE vertical loop by the reference
    # (AdjacentLoopMerging); the second section reads at a K offset what the first one writes
    with computation(PARALLEL):
        with interval(0, 1):
            b = a * 2.0
        with interval(1, None):
            c = b[0, 0, -1] + 1.0
