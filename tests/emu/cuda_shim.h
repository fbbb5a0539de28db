// CPU execution shim for the GENERATED CUDA kernels — test infrastructure only (never shipped, never
// on the product path).  The generated .cu text is compiled unchanged by g++ with -DB200_HOST_EMU and
// run lane by lane: one OS thread per lane of a warp, warp shuffles / votes through a barrier, so the
// code generator can be validated against the oracle on machines without a GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static /* warps of a CTA run one after the other, lanes of a warp are threads of this process */
#define __align__(n) __attribute__((aligned(n)))

struct emu_uint3 { unsigned x, y, z; };
extern thread_local emu_uint3 threadIdx, blockIdx;
extern emu_uint3 blockDim, gridDim;

// same alignment as the CUDA built-in vector types: the kernels are compiled with -fsanitize=alignment
// (tests/emu/emu.py), so a vector load/store the device would trap on (misaligned address) aborts here too
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) longlong2 { long long x, y; };
inline int2 make_int2(int x, int y) { return {x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline double2 make_double2(double x, double y) { return {x, y}; }
inline longlong2 make_longlong2(long long x, long long y) { return {x, y}; }

template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline double __hiloint2double(int hi, int lo) {
  unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
  double d; std::memcpy(&d, &u, 8); return d;
}
using std::isfinite; using std::isinf; using std::isnan;

namespace emu {
unsigned long long exchange(unsigned long long bits, int src_lane);  // value of `bits` held by src_lane
bool any(bool pred);
int lane();
void fail(const char* what);  // device trap / hang: abort the test run with a message
void trace(int slot);  // B200_TRACE: per-path counters, read back with emu_trace_read()
template <class T> inline T shfl(T v, int src) {
  unsigned long long b = 0; std::memcpy(&b, &v, sizeof(T));
  b = exchange(b, src);
  T r; std::memcpy(&r, &b, sizeof(T)); return r;
}
}  // namespace emu
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) { int l = emu::lane(); return emu::shfl(v, l - d >= 0 ? l - d : l); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) { int l = emu::lane(); return emu::shfl(v, l + d < 32 ? l + d : l); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl(v, src); }
inline bool __any_sync(unsigned, bool p) { return emu::any(p); }
inline void __syncwarp(unsigned = 0xffffffffu) { (void)emu::any(false); }
inline bool __all_sync(unsigned, bool p) { return !emu::any(!p); }
