// libgt4py_b200.so — the thin C-ABI launcher of the b200 stencil backend (see include/gt4py_b200.h).
//
// Stands in for the per-stencil pybind11 extension + gridtools::stencil::run of the reference
// (backend/gtc_common.py:65-103, gtc/gtcpp/gtcpp_codegen.py:267-285): loads the sm_100a cubin the
// b200 code generator produced, owns the scratch for surviving temporaries (GT_DECLARE_TMP in the
// reference, gtcpp_codegen.py:241-247), builds the kernel argument block from borrowed device
// pointers and enqueues the kernels of the launch plan on the caller's stream.
//
// Only the CUDA runtime is linked (statically); NCCL is resolved with dlopen at first use so the
// library loads on machines without NCCL/driver and fails loudly, never silently, when used there.
#include "../../include/gt4py_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? B200_ERR_NO_DEVICE \
                                                                                : B200_ERR_CUDA,   \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);     \
  } while (0)

// ---- mirror of the device-side argument block (csrc/b200_device.cuh) -----------------------------
struct FieldArg {
  char* p;
  long long s[7];
  int klo, khi;
  int vec, _pad;
};
struct Geom {
  int nI, nJ, nK;
  int i_lo, i_hi, j_lo, j_hi;
  int k_lo, k_hi;
  int _pad;
  unsigned long long* halo_flag_lo;
  unsigned long long* halo_flag_hi;
  unsigned long long halo_epoch;
};
static_assert(sizeof(FieldArg) == 80, "FieldArg layout");
static_assert(sizeof(Geom) == 64, "Geom layout");

struct Bound {
  int level;  // 0 = START, 1 = END
  int off;
  int resolve(int nK) const { return level == 0 ? off : nK + off; }
};

struct FieldPlan {
  std::string name;
  int is_temp, itemsize, dims[3], ndata, data[4];
  int ei0, ei1, ej0, ej1;
};
struct KernelPlan {
  std::string name;
  int kind;  // 0 par, 1 seq, 2 stream
  int block[3], tile[3];
  int ei0, ei1, ej0, ej1;
  Bound k_lo, k_hi;
  int smem;
  // column kernels with temporaries in shared memory (codegen_column.py, col_smem): smem_per_k bytes of dynamic shared memory
  // per level when the domain has at most smem_kcap levels (the kernel takes its global-scratch path otherwise)
  int smem_per_k = 0, smem_kcap = 0;
  int qshift = 0;  // streaming kernels: segments may start up to this many vectors left of the first stored column
  cudaKernel_t fn = nullptr;
};
struct TmapPlan {
  int field, box0, box1;
  // last encoded map (re-encoding costs ~1 us per map; calls usually repeat the same buffers)
  const void* key_base = nullptr;
  long long key[6] = {0, 0, 0, 0, 0, 0};
  alignas(64) unsigned char map[128];
};
struct SectionPlan {
  Bound k0, k1;
  std::vector<int> kernels;
};
struct StepPlan {
  int type;  // 0 launch, 1 levels
  int kernel = -1;
  int order = 0;  // 0 forward, 1 backward
  std::vector<SectionPlan> sections;
};

}  // namespace

struct b200_stencil {
  std::string name;
  std::vector<FieldPlan> fields;
  int n_api = 0;
  size_t scalars_size = 0;
  std::vector<KernelPlan> kernels;
  std::vector<StepPlan> steps;
  std::vector<TmapPlan> tmaps;
  cudaLibrary_t lib = nullptr;
  std::vector<char> image;
  // Scratch for temporaries: one grow-only buffer PER STREAM the stencil is launched on, so that calls of the same
  // stencil on different streams (interior / boundary sub-boxes overlapped with a halo exchange, device_sync=False)
  // never share temporaries.  A buffer whose address is baked into a captured CUDA graph is never freed before
  // the stencil is unloaded (it is retired when a later call needs a bigger one).
  struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
    bool captured = false;
  };
  std::map<cudaStream_t, Scratch> scratch;
  std::vector<void*> retired;
  std::mutex mu;
  int device = -1;
};

struct b200_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int n_nodes = 0;
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int parse_plan(const char* text, b200_stencil* st) {
  std::istringstream in(text);
  std::string tok;
  int version = 0;
  if (!(in >> tok >> version) || tok != "b200plan" || version != 3)
    return fail(B200_ERR_INVALID, "launch plan: bad header");
  int nfields = 0, nkernels = 0, nsteps = 0, ntmaps = 0;
  while (in >> tok) {
    if (tok == "name") {
      in >> st->name;
    } else if (tok == "nfields") {
      in >> nfields;
    } else if (tok == "field") {
      FieldPlan f;
      in >> f.name >> f.is_temp >> f.itemsize >> f.dims[0] >> f.dims[1] >> f.dims[2] >> f.ndata >> f.data[0] >>
          f.data[1] >> f.data[2] >> f.data[3] >> f.ei0 >> f.ei1 >> f.ej0 >> f.ej1;
      st->fields.push_back(f);
    } else if (tok == "scalars_size") {
      in >> st->scalars_size;
    } else if (tok == "nkernels") {
      in >> nkernels;
    } else if (tok == "kernel") {
      KernelPlan k;
      in >> k.name >> k.kind >> k.block[0] >> k.block[1] >> k.block[2] >> k.tile[0] >> k.tile[1] >> k.tile[2] >>
          k.ei0 >> k.ei1 >> k.ej0 >> k.ej1 >> k.k_lo.level >> k.k_lo.off >> k.k_hi.level >> k.k_hi.off >> k.smem >> k.qshift >> k.smem_per_k >> k.smem_kcap;
      st->kernels.push_back(k);
    } else if (tok == "ntmaps") {
      in >> ntmaps;
    } else if (tok == "tmap") {
      TmapPlan t;
      in >> t.field >> t.box0 >> t.box1;
      st->tmaps.push_back(t);
    } else if (tok == "nsteps") {
      in >> nsteps;
    } else if (tok == "step") {
      std::string kind;
      in >> kind;
      StepPlan s;
      if (kind == "launch") {
        s.type = 0;
        in >> s.kernel;
      } else if (kind == "levels") {
        s.type = 1;
        int nsec = 0;
        in >> s.order >> nsec;
        for (int i = 0; i < nsec; ++i) {
          std::string t2;
          SectionPlan sec;
          int nk = 0;
          in >> t2 >> sec.k0.level >> sec.k0.off >> sec.k1.level >> sec.k1.off >> nk;
          if (t2 != "section") return fail(B200_ERR_INVALID, "launch plan: expected 'section'");
          sec.kernels.resize(nk);
          for (int j = 0; j < nk; ++j) in >> sec.kernels[j];
          s.sections.push_back(sec);
        }
      } else {
        return fail(B200_ERR_INVALID, "launch plan: unknown step '%s'", kind.c_str());
      }
      st->steps.push_back(s);
    } else if (tok == "end") {
      break;
    } else {
      return fail(B200_ERR_INVALID, "launch plan: unknown token '%s'", tok.c_str());
    }
  }
  if (!in || (int)st->fields.size() != nfields || (int)st->kernels.size() != nkernels ||
      (int)st->steps.size() != nsteps || (int)st->tmaps.size() != ntmaps)
    return fail(B200_ERR_INVALID, "launch plan: truncated or inconsistent");
  for (auto& t : st->tmaps)
    if (t.field < 0 || t.field >= nfields || t.box0 <= 0 || t.box0 > 256 || t.box1 <= 0 || t.box1 > 256)
      return fail(B200_ERR_INVALID, "plan: tensor map");
  for (auto& f : st->fields)
    if (!f.is_temp) st->n_api++;
  for (auto& s : st->steps) {
    if (s.type == 0 && (s.kernel < 0 || s.kernel >= nkernels)) return fail(B200_ERR_INVALID, "plan: kernel index");
    for (auto& sec : s.sections)
      for (int k : sec.kernels)
        if (k < 0 || k >= nkernels) return fail(B200_ERR_INVALID, "plan: kernel index");
  }
  return B200_OK;
}

// Temporaries: I stride-1, padded to 32 elements so that the origin column is 128-byte aligned.
struct TempLayout {
  size_t offset, bytes;
  long long s[7];
  int origin[3], shape[3];
};

TempLayout temp_layout(const FieldPlan& f, const int32_t dom[3], size_t offset) {
  TempLayout t{};
  int ni = f.dims[0] ? dom[0] + (f.ei1 - f.ei0) : 1;
  int nj = f.dims[1] ? dom[1] + (f.ej1 - f.ej0) : 1;
  int nk = f.dims[2] ? dom[2] : 1;
  int lead = f.dims[0] ? (int)align_up((size_t)(-f.ei0), 32) : 0;  // pad so that origin is aligned
  long long pitch_i = f.dims[0] ? (long long)align_up((size_t)(lead + dom[0] + f.ei1), 32) : 1;
  long long nd = 1;
  for (int d = 0; d < f.ndata; ++d) nd *= f.data[d];
  // layout (fastest to slowest): I, J, K, data0, data1
  t.s[0] = f.dims[0] ? 1 : 0;
  t.s[1] = f.dims[1] ? pitch_i : 0;
  t.s[2] = f.dims[2] ? pitch_i * nj : 0;
  long long vol = pitch_i * nj * nk;
  long long acc = vol;  // data dimensions are outermost, last one fastest among them
  for (int d = f.ndata - 1; d >= 0; --d) {
    t.s[3 + d] = acc;
    acc *= f.data[d];
  }
  t.origin[0] = f.dims[0] ? lead : 0;
  t.origin[1] = f.dims[1] ? -f.ej0 : 0;
  t.origin[2] = 0;
  t.shape[0] = ni;
  t.shape[1] = nj;
  t.shape[2] = nk;
  t.offset = offset;
  t.bytes = align_up((size_t)(vol * nd) * f.itemsize, 256);
  return t;
}

// vector (16-byte) path is legal when I is unit-stride and every row start is 16-byte aligned
int vec_ok(const FieldArg& a, int itemsize) {
  if (a.s[0] != 1) return 0;
  if (((uintptr_t)a.p) & 15) return 0;
  if (((a.s[1] * itemsize) & 15) || ((a.s[2] * itemsize) & 15)) return 0;
  return 1;
}

// cuTensorMapEncodeTiled, resolved at run time (the launcher links the CUDA runtime only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// Tensor map of one field for the bulk-async streaming kernels: a 3-D tensor (I, J, K) over the WHOLE array the caller
// passed (halo included), box = box0 elements x box1 rows x 1 level.  `start` = address of array element [0,0,0]; the
// map's base is `start` rounded down to 16 bytes (always inside the same allocation: device allocations are at least
// 256-byte aligned), the skipped elements are added to the I extent and to the I origin the kernel adds to its
// coordinates.  tmo = {origin I (+ shift), origin J, origin K, K multiplier}.
int encode_tmap(TmapPlan& t, const char* start, const int shape[3], const long long s[7], const int origin[3], int itemsize,
                bool has_k, unsigned char* out_map, int tmo[4]) {
  const uintptr_t a = (uintptr_t)start;
  const char* base = start - (a & 15);
  const int extra = (int)((a & 15) / itemsize);
  const long long dimk = has_k ? shape[2] : 1;
  const long long sj = s[1] * itemsize, sk = has_k ? s[2] * itemsize : s[1] * itemsize * shape[1];
  tmo[0] = origin[0] + extra;
  tmo[1] = origin[1];
  tmo[2] = has_k ? origin[2] : 0;
  tmo[3] = has_k ? 1 : 0;
  const long long key[6] = {shape[0] + extra, shape[1], dimk, sj, sk, itemsize};
  if (t.key_base == base && memcmp(t.key, key, sizeof(key)) == 0) {
    memcpy(out_map, t.map, 128);
    return B200_OK;
  }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)key[0], (cuuint64_t)key[1], (cuuint64_t)key[2]};
  const cuuint64_t strides[2] = {(cuuint64_t)sj, (cuuint64_t)sk};
  const cuuint32_t box[3] = {(cuuint32_t)t.box0, (cuuint32_t)t.box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = itemsize == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  CUresult r = enc(&m, dt, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): dims %lld x %lld x %lld, strides %lld / %lld bytes", (int)r,
                key[0], key[1], key[2], sj, sk);
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  memcpy(t.map, &m, 128);
  t.key_base = base;
  memcpy(t.key, key, sizeof(key));
  memcpy(out_map, t.map, 128);
  return B200_OK;
}

int launch(b200_stencil* st, KernelPlan& k, std::vector<char>& blob, int k_lo, int k_hi, cudaStream_t stream) {
  Geom* g = reinterpret_cast<Geom*>(blob.data());
  int nx = (g->i_hi + k.ei1) - (g->i_lo + k.ei0);
  int ny = (g->j_hi + k.ej1) - (g->j_lo + k.ej0);
  int nz = k.kind == 1 ? 1 : (k_hi - k_lo);
  if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
  g->k_lo = k_lo;
  g->k_hi = k_hi;
  dim3 grid((nx + k.tile[0] - 1) / k.tile[0], (ny + k.tile[1] - 1) / k.tile[1], (nz + k.tile[2] - 1) / k.tile[2]);
  dim3 block(k.block[0], k.block[1], k.block[2]);
  if (k.kind == 2) {
    // streaming kernels: one warp per (I segment, J tile, K level) task, block[1] warps per CTA;
    // tile = {I points per segment, rows per tile, vector width}; segments start at the vector-aligned
    // column at or below the first stored column (must match codegen_stream.py)
    const int V = k.tile[2];
    const int x0 = g->i_lo + k.ei0, x1 = g->i_hi + k.ei1;
    const int qx0 = x0 >= 0 ? x0 / V : -((-x0 + V - 1) / V);
    // qshift > 0 (bulk-async variant: segments start on 16-byte boundaries): the kernel may shift its segment grid
    // left by up to qshift vectors; size the grid for the worst case, surplus warps leave at once
    const long long nseg = ((x1 - (qx0 - k.qshift) * V) + k.tile[0] - 1) / k.tile[0];
    const long long ntj = (ny + k.tile[1] - 1) / k.tile[1];
    const long long tasks = nseg * ntj * nz;
    grid = dim3((unsigned)((tasks + k.block[1] - 1) / k.block[1]), 1, 1);
  }
  if (grid.y > 65535 || grid.z > 65535) return fail(B200_ERR_INVALID, "grid too large for kernel %s", k.name.c_str());
  void* params[1] = {blob.data()};
  size_t smem = (size_t)k.smem;
  if (k.smem_per_k > 0 && g->nK <= k.smem_kcap) smem += (size_t)k.smem_per_k * (size_t)g->nK;
  CU(cudaLaunchKernel((const void*)k.fn, grid, block, params, smem, stream));
  return 1;
}

}  // namespace

extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }

size_t b200_sizeof_field(void) { return sizeof(b200_field_t); }

const char* b200_last_error(void) { return g_err.c_str(); }

int b200_device_info(int device, int* n_devices, int* sm_major, int* sm_minor, int* n_sms) {
  int n = 0;
  CU(cudaGetDeviceCount(&n));
  if (n_devices) *n_devices = n;
  if (device >= n) return fail(B200_ERR_NO_DEVICE, "device %d of %d", device, n);
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, device));
  if (sm_major) *sm_major = p.major;
  if (sm_minor) *sm_minor = p.minor;
  if (n_sms) *n_sms = p.multiProcessorCount;
  return B200_OK;
}

int b200_stencil_load(const void* image, size_t image_size, const char* plan_text, b200_stencil_t** out) {
  if (!image || !image_size || !plan_text || !out) return fail(B200_ERR_INVALID, "b200_stencil_load: null argument");
  b200_stencil* st = new b200_stencil();
  int rc = parse_plan(plan_text, st);
  if (rc != B200_OK) {
    delete st;
    return rc;
  }
  st->image.assign((const char*)image, (const char*)image + image_size);
  cudaError_t e = cudaGetDevice(&st->device);
  if (e == cudaSuccess) e = cudaFree(0);  // make sure the primary context exists
  if (e == cudaSuccess)
    e = cudaLibraryLoadData(&st->lib, st->image.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) {
    delete st;
    return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? B200_ERR_NO_DEVICE : B200_ERR_CUDA,
                "loading stencil image failed: %s", cudaGetErrorString(e));
  }
  for (auto& k : st->kernels) {
    e = cudaLibraryGetKernel(&k.fn, st->lib, k.name.c_str());
    const int smem_max = k.smem + k.smem_per_k * k.smem_kcap;
    if (e == cudaSuccess && smem_max > 48 * 1024)
      e = cudaFuncSetAttribute((const void*)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) {
      std::string kn = k.name;
      cudaLibraryUnload(st->lib);
      delete st;
      return fail(B200_ERR_CUDA, "kernel %s: %s", kn.c_str(), cudaGetErrorString(e));
    }
  }
  *out = st;
  return B200_OK;
}

int b200_stencil_unload(b200_stencil_t* st) {
  if (!st) return B200_OK;
  // kernels of this stencil may still be in flight (device_sync=False): wait before the code and scratch go away
  cudaDeviceSynchronize();
  for (auto& kv : st->scratch)
    if (kv.second.p) cudaFree(kv.second.p);
  for (void* p : st->retired) cudaFree(p);
  if (st->lib) cudaLibraryUnload(st->lib);
  delete st;
  return B200_OK;
}

int b200_stencil_num_fields(const b200_stencil_t* st) { return st ? st->n_api : B200_ERR_INVALID; }
size_t b200_stencil_scalars_size(const b200_stencil_t* st) { return st ? st->scalars_size : 0; }
int b200_stencil_num_kernels(const b200_stencil_t* st) { return st ? (int)st->kernels.size() : B200_ERR_INVALID; }
const char* b200_stencil_kernel_name(const b200_stencil_t* st, int index) {
  if (!st || index < 0 || index >= (int)st->kernels.size()) return nullptr;
  return st->kernels[index].name.c_str();
}

int b200_stencil_run(b200_stencil_t* st, const b200_field_t* fields, int nfields, const void* scalars,
                     size_t scalars_size, const int32_t domain[3], const int32_t subbox[4], void* stream_) {
  return b200_stencil_run_halo(st, fields, nfields, scalars, scalars_size, domain, subbox, nullptr, nullptr, 0, stream_);
}

int b200_stencil_run_halo(b200_stencil_t* st, const b200_field_t* fields, int nfields, const void* scalars,
                          size_t scalars_size, const int32_t domain[3], const int32_t subbox[4], uint64_t* flag_lo,
                          uint64_t* flag_hi, uint64_t epoch, void* stream_) {
  if (!st || !domain) return fail(B200_ERR_INVALID, "b200_stencil_run: null argument");
  if (nfields != st->n_api)
    return fail(B200_ERR_INVALID, "stencil %s expects %d fields, got %d", st->name.c_str(), st->n_api, nfields);
  if (scalars_size != st->scalars_size || (scalars_size && !scalars))
    return fail(B200_ERR_INVALID, "stencil %s expects %zu bytes of scalars, got %zu", st->name.c_str(),
                st->scalars_size, scalars_size);
  if (domain[0] <= 0 || domain[1] <= 0 || domain[2] <= 0)
    return fail(B200_ERR_INVALID, "empty compute domain (%d, %d, %d)", domain[0], domain[1], domain[2]);
  cudaStream_t stream = (cudaStream_t)stream_;

  // -- scratch for temporaries ---------------------------------------------------------------
  const size_t nf = st->fields.size();
  std::vector<TempLayout> tl(nf);
  size_t need = 0;
  for (size_t n = 0; n < nf; ++n)
    if (st->fields[n].is_temp == 1) {
      tl[n] = temp_layout(st->fields[n], domain, need);
      need += tl[n].bytes;
    }
  char* scratch_base = nullptr;
  if (need > 0) {
    std::lock_guard<std::mutex> lock(st->mu);
    b200_stencil::Scratch& sc = st->scratch[stream];
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CU(cudaStreamIsCapturing(stream, &cap));
    if (need > sc.bytes) {
      const bool capturing = cap != cudaStreamCaptureStatusNone;
      if (sc.p) {
        if (sc.captured || capturing) {
          st->retired.push_back(sc.p);  // a captured graph launches kernels that point into it: freed at unload
        } else {
          CU(cudaStreamSynchronize(stream));
          CU(cudaFree(sc.p));
        }
        sc = b200_stencil::Scratch();
      }
      // (legal while capturing: b200_graph_begin captures in relaxed mode, and nothing is freed or synchronised here)
      CU(cudaMalloc(&sc.p, need));
      sc.bytes = need;
    }
    if (cap != cudaStreamCaptureStatusNone) sc.captured = true;
    scratch_base = (char*)sc.p;
  }

  // -- argument block --------------------------------------------------------------------------
  const size_t nslots = nf ? nf : 1;
  const size_t scal_off = sizeof(Geom) + sizeof(FieldArg) * nslots;
  const size_t ntm = st->tmaps.size();
  const size_t tm_off = align_up(scal_off + st->scalars_size, 64);  // Args::tm (alignas(64)), then Args::tmo
  std::vector<char> blob(ntm ? align_up(tm_off + ntm * (128 + 16), 64) : align_up(scal_off + st->scalars_size, 8), 0);
  std::vector<const char*> arr_start(nf, nullptr);  // per field: array element [0,0,0], extents, origin (for the tensor maps)
  std::vector<std::array<int, 6>> arr_geo(nf);
  Geom* g = reinterpret_cast<Geom*>(blob.data());
  g->nI = domain[0];
  g->nJ = domain[1];
  g->nK = domain[2];
  g->i_lo = subbox ? subbox[0] : 0;
  g->i_hi = subbox ? subbox[1] : domain[0];
  g->j_lo = subbox ? subbox[2] : 0;
  g->j_hi = subbox ? subbox[3] : domain[1];
  g->halo_flag_lo = (unsigned long long*)flag_lo;
  g->halo_flag_hi = (unsigned long long*)flag_hi;
  g->halo_epoch = (flag_lo || flag_hi) ? epoch : 0;
  FieldArg* fa = reinterpret_cast<FieldArg*>(blob.data() + sizeof(Geom));
  int api = 0;
  for (size_t n = 0; n < nf; ++n) {
    const FieldPlan& fp = st->fields[n];
    FieldArg& a = fa[n];
    if (fp.is_temp == 2) {  // temporary that lives entirely in registers: no scratch
      a.p = nullptr;
      continue;
    }
    if (fp.is_temp) {
      const TempLayout& t = tl[n];
      for (int d = 0; d < 7; ++d) a.s[d] = t.s[d];
      long long off = (long long)t.origin[0] * t.s[0] + (long long)t.origin[1] * t.s[1];
      a.p = scratch_base + t.offset + off * fp.itemsize;
      a.klo = 0;
      a.khi = t.shape[2];
      a.vec = vec_ok(a, fp.itemsize);
      arr_start[n] = scratch_base + t.offset;
      // (every element of a row of the scratch layout exists: the I extent of the map is the row pitch)
      arr_geo[n] = {fp.dims[1] ? (int)t.s[1] : t.shape[0], t.shape[1], t.shape[2], t.origin[0], t.origin[1], t.origin[2]};
    } else {
      const b200_field_t& f = fields[api++];
      if (!f.data) {  // unreferenced argument (AccessKind.NONE)
        a.p = nullptr;
        continue;
      }
      long long off = 0;
      for (int d = 0; d < 7; ++d) a.s[d] = f.strides[d];
      for (int d = 0; d < 3; ++d)
        if (fp.dims[d]) off += (long long)f.origin[d] * f.strides[d];
      a.p = (char*)f.data + off * fp.itemsize;
      a.klo = fp.dims[2] ? -f.origin[2] : 0;
      a.khi = fp.dims[2] ? f.shape[2] - f.origin[2] : 1;
      a.vec = vec_ok(a, fp.itemsize);
      arr_start[n] = (const char*)f.data;
      arr_geo[n] = {f.shape[0], f.shape[1], f.shape[2], f.origin[0], f.origin[1], f.origin[2]};
    }
  }
  if (st->scalars_size) memcpy(blob.data() + scal_off, scalars, st->scalars_size);
  if (ntm) {
    std::lock_guard<std::mutex> lock(st->mu);  // the per-map encode cache
    for (size_t m = 0; m < ntm; ++m) {
      TmapPlan& t = st->tmaps[m];
      const FieldPlan& fp = st->fields[t.field];
      const FieldArg& a = fa[t.field];
      // a field off the vector path never reaches the bulk-async loop (the kernels test FieldArg::vec): map left zeroed
      if (!a.p || !a.vec || !fp.dims[0] || !fp.dims[1]) continue;
      const int shape[3] = {arr_geo[t.field][0], arr_geo[t.field][1], arr_geo[t.field][2]};
      const int origin[3] = {arr_geo[t.field][3], arr_geo[t.field][4], arr_geo[t.field][5]};
      int rc = encode_tmap(t, arr_start[t.field], shape, a.s, origin, fp.itemsize, fp.dims[2] != 0,
                           (unsigned char*)blob.data() + tm_off + 128 * m, reinterpret_cast<int*>(blob.data() + tm_off + 128 * ntm) + 4 * m);
      if (rc < 0) return rc;
    }
  }

  // -- steps -------------------------------------------------------------------------------------
  int launches = 0;
  const int nK = domain[2];
  for (auto& s : st->steps) {
    if (s.type == 0) {
      KernelPlan& k = st->kernels[s.kernel];
      int rc = launch(st, k, blob, k.k_lo.resolve(nK), k.k_hi.resolve(nK), stream);
      if (rc < 0) return rc;
      launches += rc;
    } else {
      // level-by-level: sections are listed in execution order
      for (auto& sec : s.sections) {
        int k0 = sec.k0.resolve(nK), k1 = sec.k1.resolve(nK);
        for (int kk = 0; kk < k1 - k0; ++kk) {
          int level = s.order == 0 ? k0 + kk : k1 - 1 - kk;
          for (int ki : sec.kernels) {
            int rc = launch(st, st->kernels[ki], blob, level, level + 1, stream);
            if (rc < 0) return rc;
            launches += rc;
          }
        }
      }
    }
  }
  return launches;
}

// ---- streams / events --------------------------------------------------------------------------
int b200_stream_create(void** stream) {
  cudaStream_t s;
  CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = s;
  return B200_OK;
}
int b200_stream_create_priority(void** stream, int high) {
  int least = 0, greatest = 0;  // numerically lower = higher priority
  CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  cudaStream_t s;
  CU(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high ? greatest : least));
  *stream = s;
  return B200_OK;
}
int b200_stream_destroy(void* stream) {
  CU(cudaStreamDestroy((cudaStream_t)stream));
  return B200_OK;
}
int b200_stream_synchronize(void* stream) {
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return B200_OK;
}
int b200_event_create(void** event) {
  cudaEvent_t e;
  CU(cudaEventCreate(&e));
  *event = e;
  return B200_OK;
}
int b200_event_destroy(void* event) {
  CU(cudaEventDestroy((cudaEvent_t)event));
  return B200_OK;
}
int b200_event_record(void* event, void* stream) {
  CU(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return B200_OK;
}
int b200_stream_wait_event(void* stream, void* event) {
  CU(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
  return B200_OK;
}
int b200_event_elapsed_ms(void* start, void* stop, float* ms) {
  CU(cudaEventSynchronize((cudaEvent_t)stop));
  CU(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return B200_OK;
}

// ---- stencil sequences as CUDA graphs -------------------------------------------------
// A time step is many small stencil calls (reference caller: examples/cartesian/demo_burgers.ipynb cell 12,
// 3 RK stages + copies + boundary conditions); capturing the launches the calls enqueue removes the
// per-call host overhead on replay.  Everything enqueued on `stream` (and on streams forked from it with
// b200_event_record / b200_stream_wait_event, e.g. the halo exchange) between begin and end is captured;
// kernel arguments (field addresses, scalars, domain) are frozen at capture time.
int b200_graph_begin(void* stream) {
  CU(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeRelaxed));
  return B200_OK;
}

int b200_graph_end(void* stream, b200_graph_t** out) {
  if (!out) return fail(B200_ERR_INVALID, "b200_graph_end: null argument");
  cudaGraph_t graph = nullptr;
  CU(cudaStreamEndCapture((cudaStream_t)stream, &graph));
  if (!graph) return fail(B200_ERR_CUDA, "stream capture was invalidated");
  b200_graph* g = new b200_graph();
  g->graph = graph;
  size_t n = 0;
  cudaError_t e = cudaGraphGetNodes(graph, nullptr, &n);
  if (e == cudaSuccess) g->n_nodes = (int)n;
  if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, graph, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(graph);
    delete g;
    return fail(B200_ERR_CUDA, "instantiating the captured graph failed: %s", cudaGetErrorString(e));
  }
  *out = g;
  return B200_OK;
}

int b200_graph_num_nodes(const b200_graph_t* g) { return g ? g->n_nodes : B200_ERR_INVALID; }

int b200_graph_launch(b200_graph_t* g, void* stream) {
  if (!g || !g->exec) return fail(B200_ERR_INVALID, "b200_graph_launch: null graph");
  CU(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  return B200_OK;
}

int b200_graph_destroy(b200_graph_t* g) {
  if (!g) return B200_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return B200_OK;
}

// ---- strided slab copy -----------------------------------------------------------------------------
}  // extern "C"

namespace {
// One 16-byte (or 4-byte tail-safe) lane per thread along the row, rows on blockIdx.y.
__global__ void pack2d_kernel(char* __restrict__ dst, size_t dst_pitch, const char* __restrict__ src, size_t src_pitch,
                              size_t row_bytes, size_t rows) {
  size_t row = blockIdx.y;
  size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  for (; row < rows; row += gridDim.y) {
    const char* s = src + row * src_pitch;
    char* d = dst + row * dst_pitch;
    if (x + 16 <= row_bytes && ((((size_t)s | (size_t)d) + x) & 15) == 0 && ((src_pitch | dst_pitch) & 15) == 0) {
      *reinterpret_cast<int4*>(d + x) = *reinterpret_cast<const int4*>(s + x);
    } else {
      for (size_t b = x; b < row_bytes && b < x + 16; ++b) d[b] = s[b];
    }
  }
}

// ---- NCCL through dlopen -------------------------------------------------------------------------
struct Id128 {
  char internal[B200_NCCL_UNIQUE_ID_BYTES];
};
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
  if (g_nccl.handle) return B200_OK;
  const char* names[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(B200_ERR_NCCL, "cannot load NCCL: %s", dlerror());
#define SYM(field, name)                                                        \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                    \
  if (!g_nccl.field) return fail(B200_ERR_NCCL, "NCCL symbol %s missing", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.handle = h;
  return B200_OK;
}

#define NC(call)                                                                                  \
  do {                                                                                            \
    int r_ = (call);                                                                              \
    if (r_ != 0) return fail(B200_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_));   \
  } while (0)

// Peer-memory halo push (b200_halo_push): 16-byte lanes, one grid row per (level, row) of every box; the destination
// is a peer GPU's memory mapped into this process (posted NVLink writes).
struct PushArgs {
  int n;
  b200_push_t b[16];
};
struct FlagArgs {
  int n;
  unsigned long long* f[4];
};
template <class V>
__global__ void halo_push_kernel(PushArgs a) {
  const b200_push_t& b = a.b[blockIdx.z];
  const size_t vecs = b.row_bytes / sizeof(V), total_rows = b.rows * b.levels;
  for (size_t r = blockIdx.y; r < total_rows; r += gridDim.y) {
    const size_t lev = r / b.rows, row = r % b.rows;
    const V* s = reinterpret_cast<const V*>((const char*)b.src + lev * b.src_level_pitch + row * b.src_row_pitch);
    V* d = reinterpret_cast<V*>((char*)b.dst + lev * b.dst_level_pitch + row * b.dst_row_pitch);
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < vecs; x += (size_t)gridDim.x * blockDim.x) d[x] = s[x];
  }
}
// consumer side for stencils whose kernels cannot wait themselves (point / column kernels): one thread spins on this
// rank's flags; the stencil launched behind it on the same stream starts once the neighbours' rows have landed
__global__ void halo_wait_kernel(const unsigned long long* lo, const unsigned long long* hi, unsigned long long epoch) {
  const unsigned long long* f[2] = {lo, hi};
  for (int n = 0; n < 2; ++n) {
    if (!f[n]) continue;
    unsigned long long v;
    for (long long spin = 0;; ++spin) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f[n]) : "memory");
      if (v >= epoch) break;
      __nanosleep(200);
      if (spin > 20000000LL) __trap();
    }
  }
}
__global__ void halo_flag_kernel(FlagArgs a, unsigned long long epoch) {
  if (threadIdx.x < a.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.f[threadIdx.x]), "l"(epoch) : "memory");
  }
}

// Re-layout of a 3-D array between two stride sets (storage.from_array: C-ordered upload -> pitched I-unit-stride
// storage).  When the unit-stride axes differ the copy goes through a 32 x 32 shared-memory tile so that both the
// reads and the writes are coalesced (a plain element-wise strided copy reaches ~0.4 TB/s, profiles/r01_launches_bench.csv).
struct RelayoutArgs {
  int n[3];
  long long ds[3], ss[3];
  int a, b, c, d;
};

template <class T>
__global__ void relayout_rows(T* __restrict__ dst, const T* __restrict__ src, RelayoutArgs r) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= r.n[r.a]) return;
  const long long y = blockIdx.y, z = blockIdx.z;
  dst[x * r.ds[r.a] + y * r.ds[r.c] + z * r.ds[r.d]] = src[x * r.ss[r.a] + y * r.ss[r.c] + z * r.ss[r.d]];
}

template <class T>
__global__ void relayout_tiles(T* __restrict__ dst, const T* __restrict__ src, RelayoutArgs r) {
  __shared__ T tile[32][33];
  const int a0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const long long c = blockIdx.z;
  for (int i = threadIdx.y; i < 32; i += 8) {  // read: threadIdx.x runs along the source's unit-stride axis b
    const int ia = a0 + i, ib = b0 + threadIdx.x;
    if (ia < r.n[r.a] && ib < r.n[r.b]) tile[i][threadIdx.x] = src[ia * r.ss[r.a] + ib * r.ss[r.b] + c * r.ss[r.c]];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {  // write: threadIdx.x runs along the destination's unit-stride axis a
    const int ia = a0 + threadIdx.x, ib = b0 + i;
    if (ia < r.n[r.a] && ib < r.n[r.b]) dst[ia * r.ds[r.a] + ib * r.ds[r.b] + c * r.ds[r.c]] = tile[threadIdx.x][i];
  }
}

}  // namespace

struct b200_comm {
  void* comm = nullptr;
  int n_ranks = 0, rank = 0;
};

extern "C" {

int b200_pack_2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows,
                 void* stream) {
  if (!rows || !row_bytes) return B200_OK;
  dim3 block(128);
  dim3 grid((unsigned)((row_bytes + 16 * 128 - 1) / (16 * 128)), (unsigned)(rows < 65535 ? rows : 65535));
  pack2d_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((char*)dst, dst_pitch, (const char*)src, src_pitch,
                                                          row_bytes, rows);
  CU(cudaGetLastError());
  return B200_OK;
}

// ---- peer-memory halo exchange: push my boundary rows into the neighbours' halo rows, then raise their flags -------
int b200_halo_push(const b200_push_t* boxes, int nboxes, uint64_t* const* flags, int nflags, uint64_t epoch, void* stream_) {
  if (nboxes < 0 || nflags < 0 || nboxes > 16 || nflags > 4 || (nboxes && !boxes) || (nflags && !flags))
    return fail(B200_ERR_INVALID, "b200_halo_push: bad arguments");
  cudaStream_t stream = (cudaStream_t)stream_;
  PushArgs pa;
  pa.n = nboxes;
  unsigned max_rows = 1;
  size_t max_bytes = 16, all = 0;  // OR of every address / size: the widest lane (16, 8 or 4 bytes) that divides them all
  for (int n = 0; n < nboxes; ++n) {
    const b200_push_t& b = boxes[n];
    if (!b.src || !b.dst) return fail(B200_ERR_INVALID, "b200_halo_push: null box");
    all |= (size_t)(uintptr_t)b.src | (size_t)(uintptr_t)b.dst | b.row_bytes | b.src_row_pitch | b.dst_row_pitch | b.src_level_pitch |
           b.dst_level_pitch;
    pa.b[n] = b;
    max_bytes = std::max(max_bytes, b.row_bytes);
    max_rows = std::max<unsigned>(max_rows, (unsigned)(b.rows * b.levels));
  }
  const int lane = (all & 15) == 0 ? 16 : ((all & 7) == 0 ? 8 : ((all & 3) == 0 ? 4 : 0));
  if (nboxes && !lane) return fail(B200_ERR_INVALID, "b200_halo_push: rows must be multiples of 4 bytes at 4-byte aligned addresses");
  if (nboxes) {
    const unsigned chunks = (unsigned)std::min<size_t>((max_bytes / lane + 127) / 128, 64);
    dim3 grid(chunks, std::min<unsigned>(max_rows, 65535u), (unsigned)nboxes);
    if (lane == 16) halo_push_kernel<int4><<<grid, 128, 0, stream>>>(pa);
    else if (lane == 8) halo_push_kernel<int2><<<grid, 128, 0, stream>>>(pa);
    else halo_push_kernel<int><<<grid, 128, 0, stream>>>(pa);
    CU(cudaGetLastError());
  }
  if (nflags) {
    FlagArgs fa;
    fa.n = nflags;
    for (int n = 0; n < nflags; ++n) fa.f[n] = (unsigned long long*)flags[n];
    // a second launch on the same stream: every byte of the push kernel has been written before this one starts
    halo_flag_kernel<<<1, 32, 0, stream>>>(fa, epoch);
    CU(cudaGetLastError());
  }
  return B200_OK;
}

int b200_halo_wait(const uint64_t* flag_lo, const uint64_t* flag_hi, uint64_t epoch, void* stream) {
  if (!flag_lo && !flag_hi) return B200_OK;
  halo_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const unsigned long long*)flag_lo, (const unsigned long long*)flag_hi, epoch);
  CU(cudaGetLastError());
  return B200_OK;
}

// ---- data movement helpers of the storage / host-call path -------------------------------------------------------
int b200_copy_box(void* dst, size_t dst_pitch, size_t dst_level_rows, const void* src, size_t src_pitch,
                  size_t src_level_rows, size_t row_bytes, size_t rows, size_t levels, void* stream) {
  if (!row_bytes || !rows || !levels) return B200_OK;
  if (!dst || !src || row_bytes > dst_pitch || row_bytes > src_pitch || rows > dst_level_rows || rows > src_level_rows)
    return fail(B200_ERR_INVALID, "b200_copy_box: box does not fit the pitched buffers");
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src), src_pitch, row_bytes, src_level_rows);
  p.dstPtr = make_cudaPitchedPtr(dst, dst_pitch, row_bytes, dst_level_rows);
  p.extent = make_cudaExtent(row_bytes, rows, levels);
  p.kind = cudaMemcpyDefault;
  CU(cudaMemcpy3DAsync(&p, (cudaStream_t)stream));
  return B200_OK;
}

int b200_relayout(void* dst, const void* src, int itemsize, const int32_t shape[3], const int64_t dst_strides[3],
                  const int64_t src_strides[3], void* stream) {
  if (!dst || !src || !shape || !dst_strides || !src_strides) return fail(B200_ERR_INVALID, "b200_relayout: null argument");
  for (int d = 0; d < 3; ++d)
    if (shape[d] <= 0) return B200_OK;
  // a = unit-stride axis of the destination, b = unit-stride axis of the source (fall back to the smallest stride)
  // (among the axes that have more than one element)
  int a = -1, b = -1;
  for (int d = 0; d < 3; ++d) {
    if (shape[d] == 1) continue;
    if (a < 0 || llabs(dst_strides[d]) < llabs(dst_strides[a])) a = d;
    if (b < 0 || llabs(src_strides[d]) < llabs(src_strides[b])) b = d;
  }
  if (a < 0) a = b = 0;  // a single element
  RelayoutArgs r;
  for (int d = 0; d < 3; ++d) {
    r.n[d] = shape[d];
    r.ds[d] = dst_strides[d];
    r.ss[d] = src_strides[d];
  }
  r.a = a;
  r.b = b;
  cudaStream_t st = (cudaStream_t)stream;
  if (a == b) {
    // same fast axis: rows along a, coalesced on both sides
    int o0 = (a + 1) % 3, o1 = (a + 2) % 3;
    r.c = o0;
    r.d = o1;
    dim3 block(256), grid((unsigned)((shape[a] + 255) / 256), (unsigned)shape[o0], (unsigned)shape[o1]);
    if (grid.y > 65535 || grid.z > 65535) return fail(B200_ERR_INVALID, "b200_relayout: extent too large");
    switch (itemsize) {
      case 1: relayout_rows<uint8_t><<<grid, block, 0, st>>>((uint8_t*)dst, (const uint8_t*)src, r); break;
      case 2: relayout_rows<uint16_t><<<grid, block, 0, st>>>((uint16_t*)dst, (const uint16_t*)src, r); break;
      case 4: relayout_rows<uint32_t><<<grid, block, 0, st>>>((uint32_t*)dst, (const uint32_t*)src, r); break;
      case 8: relayout_rows<uint64_t><<<grid, block, 0, st>>>((uint64_t*)dst, (const uint64_t*)src, r); break;
      default: return fail(B200_ERR_INVALID, "b200_relayout: item size %d", itemsize);
    }
  } else {
    r.c = 3 - a - b;
    r.d = -1;
    dim3 block(32, 8), grid((unsigned)((shape[a] + 31) / 32), (unsigned)((shape[b] + 31) / 32), (unsigned)shape[r.c]);
    if (grid.y > 65535 || grid.z > 65535) return fail(B200_ERR_INVALID, "b200_relayout: extent too large");
    switch (itemsize) {
      case 1: relayout_tiles<uint8_t><<<grid, block, 0, st>>>((uint8_t*)dst, (const uint8_t*)src, r); break;
      case 2: relayout_tiles<uint16_t><<<grid, block, 0, st>>>((uint16_t*)dst, (const uint16_t*)src, r); break;
      case 4: relayout_tiles<uint32_t><<<grid, block, 0, st>>>((uint32_t*)dst, (const uint32_t*)src, r); break;
      case 8: relayout_tiles<uint64_t><<<grid, block, 0, st>>>((uint64_t*)dst, (const uint64_t*)src, r); break;
      default: return fail(B200_ERR_INVALID, "b200_relayout: item size %d", itemsize);
    }
  }
  CU(cudaGetLastError());
  return B200_OK;
}

int b200_comm_unique_id(void* id_out) {
  int rc = nccl_load();
  if (rc) return rc;
  NC(g_nccl.GetUniqueId(id_out));
  return B200_OK;
}

int b200_comm_init(b200_comm_t** out, const void* unique_id, int n_ranks, int rank) {
  int rc = nccl_load();
  if (rc) return rc;
  Id128 id;
  memcpy(&id, unique_id, sizeof id);
  b200_comm* c = new b200_comm();
  c->n_ranks = n_ranks;
  c->rank = rank;
  int r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != 0) {
    delete c;
    return fail(B200_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
  }
  *out = c;
  return B200_OK;
}

int b200_comm_destroy(b200_comm_t* comm) {
  if (!comm) return B200_OK;
  if (comm->comm) g_nccl.CommDestroy(comm->comm);
  delete comm;
  return B200_OK;
}

int b200_halo_exchange(b200_comm_t* comm, const b200_halo_t* halos, int n_halos, int peer_lo, int peer_hi,
                       void* stream_) {
  if (!comm || (!halos && n_halos)) return fail(B200_ERR_INVALID, "b200_halo_exchange: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int ncclChar = 0;
  NC(g_nccl.GroupStart());
  for (int n = 0; n < n_halos; ++n) {
    const b200_halo_t& h = halos[n];
    if (peer_lo >= 0) {
      NC(g_nccl.Send(h.send_lo, h.bytes, ncclChar, peer_lo, comm->comm, stream));
      NC(g_nccl.Recv(h.recv_lo, h.bytes, ncclChar, peer_lo, comm->comm, stream));
    }
    if (peer_hi >= 0) {
      NC(g_nccl.Send(h.send_hi, h.bytes, ncclChar, peer_hi, comm->comm, stream));
      NC(g_nccl.Recv(h.recv_hi, h.bytes, ncclChar, peer_hi, comm->comm, stream));
    }
  }
  NC(g_nccl.GroupEnd());
  return B200_OK;
}

}  // extern "C"
