"""b200::div_inv (division by a launch-invariant divisor: Markstein's sequence on the hoisted reciprocal, guarded by
exponent-range tests, csrc/b200_device.cuh) must be bit-for-bit the IEEE quotient for EVERY operand: the header is
compiled for the host (the same code the CPU emulator runs) and compared with `a / d` over all 2^23 significands of the
dividend for a set of divisors (all-ones / power-of-two / random significands), over random bit patterns, and over the
special values (zeros of both signs, subnormals, infinities, NaNs, operands whose quotient under- or overflows)."""
import pathlib
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent

SRC = r"""
#include "cuda_shim.h"
#include "b200_device.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
template <class T, class U> static bool same(T a, T b) { U x, y; std::memcpy(&x, &a, sizeof(T)); std::memcpy(&y, &b, sizeof(T)); return x == y || (a != a && b != b); }
static float u2f(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static double u2d(unsigned long long u) { double f; std::memcpy(&f, &u, 8); return f; }
int main() {
  std::mt19937_64 rng(7);
  long bad = 0, n = 0, fast = 0;
  const float specials[] = {0.0f, -0.0f, 1.0f, -1.0f, u2f(1), u2f(0x007fffff), u2f(0x00800000), u2f(0x7f7fffff), u2f(0x7f800000), u2f(0xff800000), u2f(0x7fc00000),
                            1e-38f, -1e-38f, 1e38f, 3.0f, 1e-20f, 1e20f, u2f(0x3f7fffff), u2f(0x3fffffff)};
  // 1. every significand of the dividend, a few exponents, a set of divisors
  for (int di = 0; di < 24; ++di) {
    unsigned dm = di == 0 ? 0x7fffff : di == 1 ? 0 : di == 2 ? 1 : di == 3 ? 0x7ffffe : (unsigned)(rng() & 0x7fffff);
    unsigned de = 127 + (int)(rng() % 50) - 25;
    float d = u2f((de << 23) | dm);
    if (di & 1) d = -d;
    const auto dv = b200::div_inv_make(d);
    for (int ex : {127, 90, 170}) {
      for (unsigned m = 0; m < (1u << 23); m += (ex == 127 ? 1 : 5)) {
        const float a = u2f(((unsigned)ex << 23) | m);
        unsigned acc = 0; (void)b200::div_inv_try(a, dv, acc); fast += !b200::div_inv_bad(acc);
        ++n; if (!same<float, unsigned>(b200::div_inv(a, dv), a / d)) { if (bad++ < 5) std::printf("f32 a=%a d=%a\n", a, d); }
      }
    }
  }
  // 2. random bit patterns (all exponents, both operands) + specials
  for (long it = 0; it < 20000000; ++it) {
    float a = u2f((unsigned)rng()), d = u2f((unsigned)rng());
    if (it < 19 * 19) { a = specials[it % 19]; d = specials[it / 19]; }
    const auto dv = b200::div_inv_make(d);
    ++n; if (!same<float, unsigned>(b200::div_inv(a, dv), a / d)) { if (bad++ < 5) std::printf("f32 a=%a d=%a\n", a, d); }
  }
  const double dspecials[] = {0.0, -0.0, 1.0, -1.0, u2d(1), u2d(0x000fffffffffffffULL), u2d(0x0010000000000000ULL), u2d(0x7fefffffffffffffULL), u2d(0x7ff0000000000000ULL),
                              u2d(0xfff0000000000000ULL), u2d(0x7ff8000000000000ULL), 1e-308, 1e308, 3.0, 1e-200, 1e200, u2d(0x3fefffffffffffffULL), u2d(0x3fffffffffffffffULL), -7.0};
  for (long it = 0; it < 30000000; ++it) {
    double a, d;
    if (it < 19 * 19) { a = dspecials[it % 19]; d = dspecials[it / 19]; }
    else if (it & 1) { a = u2d(rng()); d = u2d(rng()); }
    else {  // moderate exponents: the fast path
      a = u2d((rng() & 0x800fffffffffffffULL) | ((unsigned long long)(1023 + (int)(rng() % 400) - 200) << 52));
      d = u2d((rng() & 0x800fffffffffffffULL) | ((unsigned long long)(1023 + (int)(rng() % 300) - 150) << 52));
      if ((it & 7) == 2) d = u2d((rng() & 0x8000000000000000ULL) | 0x000fffffffffffffULL | ((unsigned long long)(1023 + (int)(rng() % 300) - 150) << 52));
    }
    const auto dv = b200::div_inv_make(d);
    unsigned long long acc = 0; (void)b200::div_inv_try(a, dv, acc); fast += !b200::div_inv_bad(acc);
    ++n; if (!same<double, unsigned long long>(b200::div_inv(a, dv), a / d)) { if (bad++ < 10) std::printf("f64 a=%a d=%a\n", a, d); }
  }
  std::printf("checked %ld fast %ld bad %ld\n", n, fast, bad);
  return bad != 0;
}
"""


def test_div_inv_is_the_ieee_quotient(tmp_path):
    src = tmp_path / "div_inv.cpp"
    src.write_text(SRC)
    exe = tmp_path / "div_inv"
    stub = tmp_path / "emu_stub.cpp"
    stub.write_text(
        '#include "cuda_shim.h"\nthread_local emu_uint3 threadIdx, blockIdx; emu_uint3 blockDim, gridDim;\n'
        "namespace emu { unsigned long long exchange(unsigned long long b, int) { return b; } bool any(bool p) { return p; } int lane() { return 0; }\n"
        "void fail(const char*) { std::abort(); } void trace(int) {} }\n"
    )
    cmd = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-DB200_HOST_EMU", "-I", str(ROOT / "tests" / "emu"), "-I", str(ROOT / "gt4py_b200" / "csrc"),
           str(src), str(stub), "-o", str(exe)]  # fmt: skip
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:]
    checked, fast = (int(run.stdout.split()[n]) for n in (1, 3))
    assert checked > 60_000_000 and fast > 0.5 * checked  # the fast path is what was tested, not only the fallback
