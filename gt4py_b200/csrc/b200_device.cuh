// Device-side runtime of the b200 stencil backend: argument block shared by every generated kernel
// and the arithmetic helpers that give GTScript native functions NumPy-ufunc semantics
// (reference oracle arithmetic: src/gt4py/cartesian/gtc/ufuncs.py:15-93, gtc/common.py:910-993).
#pragma once
#ifdef B200_HOST_EMU
#include "cuda_shim.h"  // tests/emu: executes the generated kernels on the CPU (test infrastructure only)
#else
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>

// Path markers of the generated kernels: nothing on the device; counted by the CPU emulator so the
// tests can assert WHICH loop of a kernel (steady / edge / general) a case actually went through.
#ifdef B200_HOST_EMU
#define B200_TRACE(slot) emu::trace(slot)
#else
#define B200_TRACE(slot) ((void)0)
#endif

namespace b200 {

// One API field or backend-owned temporary, as seen by a kernel.
//   p      : address of the element at the *origin* (domain point (0,0,0), data index 0)
//   s[0..2]: element strides along I, J, K (0 when the field lacks the axis -> broadcast)
//   s[3..6]: element strides of up to four data dimensions
//   klo/khi: valid K index range relative to the origin (for clipping variable-K accesses,
//            reference: cartesian/utils/field.py:54-58)
//   vec    : 1 when I is unit-stride and origin / J / K strides are 16-byte aligned (vector path)
struct FieldArg {
  char* p;
  long long s[7];
  int klo, khi;
  int vec, _pad;
};
static_assert(sizeof(FieldArg) == 80, "FieldArg layout (mirrored in launcher.cu, tests/emu/emu.py)");

// Launch geometry common to all kernels of a stencil call.
struct Geom {
  int nI, nJ, nK;    // compute domain
  int i_lo, i_hi;    // horizontal sub-box of the domain this launch covers (domain coordinates);
  int j_lo, j_hi;    //   every stage additionally extends it by its own block extent
  int k_lo, k_hi;    // K range override for level-by-level launches (k_lo < 0: use the section intervals)
  int _pad;
  // Multi-GPU J-slab runs with the peer-memory halo exchange (b200_halo_push): the neighbours store their boundary rows
  // straight into this rank's halo rows over NVLink and then set these flags (in this rank's memory) to the step number.
  // Kernels generated with `halo_wait` make the tiles that read halo rows wait for flag >= halo_epoch; 0 = no waiting.
  unsigned long long* halo_flag_lo;
  unsigned long long* halo_flag_hi;
  unsigned long long halo_epoch;
};
static_assert(sizeof(Geom) == 64, "Geom layout (mirrored in launcher.cu, tests/emu/emu.py)");

template <class T>
__device__ __forceinline__ T ld(const T* p) {
  return *p;
}
template <class T>
__device__ __forceinline__ T ldro(const T* p) {  // read-only path (field not written by this kernel)
  return __ldg(p);
}
template <>
__device__ __forceinline__ bool ldro<bool>(const bool* p) {
  return __ldg(reinterpret_cast<const unsigned char*>(p)) != 0;
}

// fire-and-forget prefetch of the line holding `p` into L2 (SASS: CCTL.E.PF2)
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef B200_HOST_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// ---- bulk-async (TMA engine) row copies global -> shared, completion on an mbarrier ---------------
// Used by the `tma` variant of the streaming kernels (codegen_stream.py): one elected lane per warp issues
// `cp.async.bulk` copies of whole warp rows into a per-warp ring in shared memory D march steps ahead; the
// lanes wait on the slot's mbarrier and read their vectors with LDS.  SASS: UBLKCP.S.G + SYNCS.
// A 64-bit word in shared memory per barrier.  `mbar_wait` traps instead of hanging if a phase never completes
// (protocol error): a hung box would cost the GPU session, a trap is a loud launch failure.
#ifndef B200_HOST_EMU
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// generic-proxy writes to global memory (made visible to this thread by an acquire) -> later async-proxy reads (bulk copies)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// src, dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// One (field, box) tensor map of a kernel: 128 opaque bytes encoded by the launcher (cuTensorMapEncodeTiled) into the
// kernel's argument block.  tma_load_3d copies the box whose first element is (c0, c1, c2) in ARRAY index space
// into shared memory (rows packed back to back; elements outside the array read as zero).  SASS: UTMALDG.3D.
struct alignas(64) TMap {
  unsigned char b[128];
};
__device__ __forceinline__ void tma_load_3d(void* dst, const TMap* tm, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// true in exactly one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)  // suspend-time hint: at most ~20 us asleep per try
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  for (int spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1 << 15)) __trap();  // ~0.5 s of failed tries: a protocol error, not a slow copy
}
#else
// CPU emulation (tests/emu): copies complete at issue; a barrier word counts completed phases (low half) and
// pending transaction bytes (high half), so a wait on the wrong parity / an overrun producer aborts the test.
__device__ __forceinline__ bool elect_one() { return emu::lane() == 0; }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void fence_async_smem() {}
__device__ __forceinline__ void fence_proxy_async_global() {}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  if ((*bar >> 32) != 0) emu::fail("mbarrier: expect_tx on a barrier with a phase still in flight");
  *bar += (unsigned long long)bytes << 32;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (bytes & 15) || bytes == 0) emu::fail("cp.async.bulk: misaligned address or size");
  if ((*bar >> 32) < bytes) emu::fail("cp.async.bulk: more bytes than the barrier expects");
  memcpy(dst, src, bytes);
  *bar -= (unsigned long long)bytes << 32;
  if ((*bar >> 32) == 0) *bar += 1;  // phase complete
}
// emulated tensor map (filled by tests/emu/emu.py the way launcher.cu fills the real one)
struct alignas(64) TMap {
  union {
    unsigned char b[128];
    struct {
      char* base;
      long long dim[3], stride[3];  // extents in elements, strides in bytes (stride[0] = element size)
      int box[2], isz;
    } e;
  };
};
__device__ __forceinline__ void tma_load_3d(void* dst, const TMap* tm, int c0, int c1, int c2, unsigned long long* bar) {
  const auto& m = tm->e;
  if (!m.base) emu::fail("tensor map was not encoded");
  if (((uintptr_t)dst & 127) || ((uintptr_t)m.base & 15) || (m.stride[1] & 15) || (m.stride[2] & 15))
    emu::fail("cp.async.bulk.tensor: misaligned shared destination, global base or stride");
  if (((uintptr_t)(m.base + (long long)c0 * m.isz)) & 15)
    emu::fail("cp.async.bulk.tensor: the box must start on a 16-byte boundary along the unit-stride axis (illegal instruction on the device)");
  const unsigned bytes = (unsigned)(m.box[0] * m.box[1] * m.isz);
  if ((*bar >> 32) < bytes) emu::fail("cp.async.bulk.tensor: more bytes than the barrier expects");
  char* out = (char*)dst;
  for (int r = 0; r < m.box[1]; ++r)
    for (int x = 0; x < m.box[0]; ++x, out += m.isz) {
      const long long i = c0 + x, j = c1 + r, k = c2;
      if (i < 0 || i >= m.dim[0] || j < 0 || j >= m.dim[1] || k < 0 || k >= m.dim[2]) memset(out, 0, m.isz);
      else memcpy(out, m.base + i * m.isz + j * m.stride[1] + k * m.stride[2], m.isz);
    }
  *bar -= (unsigned long long)bytes << 32;
  if ((*bar >> 32) == 0) *bar += 1;  // phase complete
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  __syncwarp();  // the issuing lane has run everything it issued before this point
  const unsigned long long v = *bar;
  if ((v >> 32) != 0) emu::fail("mbarrier: wait would hang (transaction bytes outstanding)");
  const unsigned phases = (unsigned)v;
  if (phases == 0 || ((phases - 1) & 1) != parity) emu::fail("mbarrier: wait on the wrong phase parity (hang or overrun on the device)");
  __syncwarp();
}
#endif

// wait until a peer has stored a value >= `epoch` into `*flag` (system-scope acquire: the halo rows the peer wrote before
// the flag are visible afterwards).  Traps instead of hanging forever if the peer never arrives.
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch) {
#ifndef B200_HOST_EMU
  unsigned long long v;
  for (long long spin = 0;; ++spin) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= epoch) break;
    __nanosleep(200);
    if (spin > 20000000LL) __trap();  // ~ several seconds
  }
#else
  if (*flag < epoch) emu::fail("halo flag wait would hang (the peer has not pushed this step)");
#endif
}

// the same wait without the time limit (kernels generated with `halo_lean`): ANY second exit of the spin loop — counter +
// trap, bounded loop, out-of-line trap — costs the surrounding kernel 13-24 registers (hdiff: 40 -> 53-64, measured with
// cuobjdump); a peer that never arrives then hangs the kernel until the host's watchdog ends the process.
__device__ __forceinline__ void wait_flag_nolimit(const unsigned long long* flag, unsigned long long epoch) {
#ifndef B200_HOST_EMU
  unsigned long long v;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= epoch) break;
    __nanosleep(200);
  }
#else
  if (*flag < epoch) emu::fail("halo flag wait would hang (the peer has not pushed this step)");
#endif
}

__device__ __forceinline__ int clampk(long long k, int lo, int hi) {
  return (int)(k < lo ? lo : (k > hi - 1 ? hi - 1 : k));
}

// dynamic shared memory of a kernel (sized by the launcher); the CPU emulation runs the threads of a CTA one after the
// other, each through the whole kernel, with a static buffer
#ifndef B200_HOST_EMU
#define B200_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#define B200_DYN_SMEM(name) static unsigned char name[232448] __attribute__((aligned(16)))
#endif

// ---- division by a loop-invariant divisor -----------------------------------------------------------
// a / d where d is the same for every cell of the launch (an expression of scalar parameters and literals: grid
// spacings, time steps): the generators hoist d and its correctly rounded reciprocal r = 1 / d out of the march loop
// (DivInv, one IEEE division per thread) and every cell pays three FP instructions instead of the ~10-instruction
// division sequence.  Markstein's theorem: with r = RN(1/d), q0 = RN(a r), e = a - q0 d (exact in one FMA), the value
// q = RN(q0 + e r) IS the correctly rounded quotient RN(a / d) as long as nothing under- or overflows on the way; the
// exponent-range guards below (on d once, on the result per cell) keep every intermediate normal, anything else — zeros,
// infinities, NaNs, tiny or huge operands — takes the IEEE division.  Bit-for-bit equal to a / d: checked exhaustively
// over all 2^23 significands for hundreds of divisors (tests/test_div_inv.py), and by every parity test that divides.
template <class T>
struct DivInv {
  T d, r;   // r = RN(1 / d), or NaN when |d| is outside the guarded exponent range (every quotient then takes the IEEE path)
};
__device__ __noinline__ float div_ieee(float a, float d) { return a / d; }
__device__ __noinline__ double div_ieee(double a, double d) { return a / d; }
#ifdef B200_HOST_EMU
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline unsigned __float_as_uint(float v) { unsigned u; std::memcpy(&u, &v, 4); return u; }
inline long long __double_as_longlong(double v) { long long u; std::memcpy(&u, &v, 8); return u; }
#endif
__device__ __forceinline__ DivInv<float> div_inv_make(float d) {
  const unsigned m = __float_as_uint(d) * 2u - 0x61000000u;   // |d| in [2^-30, 2^30): exponent field, sign shifted out
  return {d, m < 0x3c000000u ? 1.0f / d : __int_as_float(0x7fc00000)};
}
__device__ __forceinline__ DivInv<double> div_inv_make(double d) {
  const unsigned long long m = (unsigned long long)__double_as_longlong(d) * 2ULL - 0x66e0000000000000ULL;   // |d| in [2^-200, 2^200)
  return {d, m < 0x3200000000000000ULL ? 1.0 / d : __longlong_as_double(0x7ff8000000000000LL)};
}
// The quotient by Markstein's sequence; `bad` accumulates (unsigned max) the exponent-range test of every quotient of
// a group, so that ONE compare + branch guards the whole group: if it fires, the group is redone with IEEE divisions.
__device__ __forceinline__ float div_inv_try(float a, const DivInv<float>& v, unsigned& bad) {
  const float q0 = __fmul_rn(a, v.r);
  const float e = __fmaf_rn(-q0, v.d, a);
  const float q = __fmaf_rn(e, v.r, q0);
  const unsigned m = __float_as_uint(q) * 2u - 0x43000000u;   // < 0x78000000: |q| in [2^-60, 2^60) (a NaN, an infinity, a zero are outside)
  bad = bad > m ? bad : m;
  return q;
}
__device__ __forceinline__ bool div_inv_bad(unsigned bad) { return bad >= 0x78000000u; }
__device__ __forceinline__ double div_inv_try(double a, const DivInv<double>& v, unsigned long long& bad) {
  const double q0 = __dmul_rn(a, v.r);
  const double e = __fma_rn(-q0, v.d, a);
  const double q = __fma_rn(e, v.r, q0);
  const unsigned long long m = (unsigned long long)__double_as_longlong(q) * 2ULL - 0x4160000000000000ULL;   // |q| in [2^-500, 2^500)
  bad = bad > m ? bad : m;
  return q;
}
__device__ __forceinline__ bool div_inv_bad(unsigned long long bad) { return bad >= 0x7d00000000000000ULL; }
// single quotient (generators that do not group)
__device__ __forceinline__ float div_inv(float a, const DivInv<float>& v) {
  unsigned bad = 0u;
  const float q = div_inv_try(a, v, bad);
  return div_inv_bad(bad) ? div_ieee(a, v.d) : q;
}
__device__ __forceinline__ double div_inv(double a, const DivInv<double>& v) {
  unsigned long long bad = 0ULL;
  const double q = div_inv_try(a, v, bad);
  return div_inv_bad(bad) ? div_ieee(a, v.d) : q;
}

// ---- NumPy-semantics helpers -------------------------------------------------------------------
// np.minimum / np.maximum propagate NaN (fmin/fmax do not).
__device__ __forceinline__ float min_(float a, float b) { return (isnan(a) || a < b) ? a : b; }
__device__ __forceinline__ double min_(double a, double b) { return (isnan(a) || a < b) ? a : b; }
__device__ __forceinline__ float max_(float a, float b) { return (isnan(a) || a > b) ? a : b; }
__device__ __forceinline__ double max_(double a, double b) { return (isnan(a) || a > b) ? a : b; }
template <class T>
__device__ __forceinline__ T min_(T a, T b) { return a < b ? a : b; }
template <class T>
__device__ __forceinline__ T max_(T a, T b) { return a > b ? a : b; }

__device__ __forceinline__ float abs_(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_(double a) { return fabs(a); }
template <class T>
__device__ __forceinline__ T abs_(T a) { return a < 0 ? (T)(-a) : a; }
template <>
__device__ __forceinline__ bool abs_<bool>(bool a) { return a; }

// np.remainder: result has the sign of the divisor (Python %), unlike C fmod.
__device__ __forceinline__ float mod_(float a, float b) {
  float r = fmodf(a, b);
  if (r != 0.0f && ((r < 0.0f) != (b < 0.0f))) r += b;
  else if (r == 0.0f) r = copysignf(0.0f, b);
  return r;
}
__device__ __forceinline__ double mod_(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
  else if (r == 0.0) r = copysign(0.0, b);
  return r;
}
template <class T>
__device__ __forceinline__ T mod_(T a, T b) {
  if (b == 0) return 0;
  T r = a % b;
  if (r != 0 && ((r < 0) != (b < 0))) r += b;
  return r;
}

__device__ __forceinline__ float pow_(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double pow_(double a, double b) { return pow(a, b); }
__device__ __forceinline__ double pow_(double a, float b) { return pow(a, (double)b); }
__device__ __forceinline__ double pow_(float a, double b) { return pow((double)a, b); }
template <class T, class U>
__device__ __forceinline__ T pow_(T a, U b) {  // integer power (np.power on ints; negative exponent -> 0 here)
  T r = 1;
  long long e = (long long)b;
  if (e < 0) return (a == 1) ? (T)1 : (T)0;
  while (e) {
    if (e & 1) r = (T)(r * a);
    a = (T)(a * a);
    e >>= 1;
  }
  return r;
}
__device__ __forceinline__ float pow_(float a, int b) { return powf(a, (float)b); }
__device__ __forceinline__ float pow_(float a, long long b) { return powf(a, (float)b); }
__device__ __forceinline__ double pow_(double a, int b) { return pow(a, (double)b); }
__device__ __forceinline__ double pow_(double a, long long b) { return pow(a, (double)b); }

// np.round == rint (half to even); round_away_from_zero = copysign(floor(|x| + 0.5), x)
__device__ __forceinline__ float round_(float a) { return rintf(a); }
__device__ __forceinline__ double round_(double a) { return rint(a); }
__device__ __forceinline__ float round_away_(float a) { return copysignf(floorf(fabsf(a) + 0.5f), a); }
__device__ __forceinline__ double round_away_(double a) { return copysign(floor(fabs(a) + 0.5), a); }

#define B200_UNARY(NAME, F32FN, F64FN)                                        \
  __device__ __forceinline__ float NAME(float a) { return F32FN(a); }         \
  __device__ __forceinline__ double NAME(double a) { return F64FN(a); }
B200_UNARY(sin_, sinf, sin)
B200_UNARY(cos_, cosf, cos)
B200_UNARY(tan_, tanf, tan)
B200_UNARY(asin_, asinf, asin)
B200_UNARY(acos_, acosf, acos)
B200_UNARY(atan_, atanf, atan)
B200_UNARY(sinh_, sinhf, sinh)
B200_UNARY(cosh_, coshf, cosh)
B200_UNARY(tanh_, tanhf, tanh)
B200_UNARY(asinh_, asinhf, asinh)
B200_UNARY(acosh_, acoshf, acosh)
B200_UNARY(atanh_, atanhf, atanh)
B200_UNARY(sqrt_, sqrtf, sqrt)
B200_UNARY(exp_, expf, exp)
B200_UNARY(log_, logf, log)
B200_UNARY(log10_, log10f, log10)
B200_UNARY(gamma_, tgammaf, tgamma)
B200_UNARY(cbrt_, cbrtf, cbrt)
B200_UNARY(floor_, floorf, floor)
B200_UNARY(ceil_, ceilf, ceil)
B200_UNARY(trunc_, truncf, trunc)
B200_UNARY(erf_, erff, erf)
B200_UNARY(erfc_, erfcf, erfc)
#undef B200_UNARY
// integer arguments of float-only functions behave like NumPy: promote to double
template <class T> __device__ __forceinline__ double sqrt_(T a) { return sqrt((double)a); }
template <class T> __device__ __forceinline__ T floor_(T a) { return a; }
template <class T> __device__ __forceinline__ T ceil_(T a) { return a; }
template <class T> __device__ __forceinline__ T trunc_(T a) { return a; }

__device__ __forceinline__ bool isfinite_(float a) { return isfinite(a); }
__device__ __forceinline__ bool isfinite_(double a) { return isfinite(a); }
__device__ __forceinline__ bool isinf_(float a) { return isinf(a); }
__device__ __forceinline__ bool isinf_(double a) { return isinf(a); }
__device__ __forceinline__ bool isnan_(float a) { return isnan(a); }
__device__ __forceinline__ bool isnan_(double a) { return isnan(a); }
template <class T> __device__ __forceinline__ bool isfinite_(T) { return true; }
template <class T> __device__ __forceinline__ bool isinf_(T) { return false; }
template <class T> __device__ __forceinline__ bool isnan_(T) { return false; }

}  // namespace b200
