/* gt4py_b200.h — C ABI of the b200 stencil-execution backend (libgt4py_b200.so).
 *
 * This is the drop-in boundary of the hot path: the native entry the reference reaches from the
 * generated module's `run()` (reference: src/gt4py/cartesian/backend/templates/stencil_module.py.in:160-169).
 * For `gt:gpu` that entry is a per-stencil pybind11 function
 *     run_computation(std::array<uint,3> domain, {py::object field, std::array<int,ndim> origin}…,
 *                     scalars…, py::object exec_info)
 * (reference: backend/gtc_common.py:65-103, 144-168; backend/gtcpp_backend.py:77-99) which reads the
 * device buffers through __cuda_array_interface__ (gt::as_cuda_sid, gtc_common.py:42-50), shifts
 * them to the origin and calls gridtools::stencil::run(spec, gpu<>{}, grid, fields…)
 * (gtc/gtcpp/gtcpp_codegen.py:267-285).  Here the same information crosses a plain C boundary:
 * pointers, strides, origins, a domain and a blob of scalar parameters.  No Python, torch or CUDA
 * types appear in the signatures; `stream` is a cudaStream_t passed as void*.
 *
 * All functions return 0 on success or a negative b200_status; the message of the last failure on
 * the calling thread is available from b200_last_error().  Nothing here falls back to the CPU.
 */
#ifndef GT4PY_B200_H
#define GT4PY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1
#define B200_MAX_DIMS 7 /* I, J, K + up to four data dimensions */

typedef enum b200_status {
  B200_OK = 0,
  B200_ERR_INVALID = -1,  /* bad argument / malformed plan                                  */
  B200_ERR_CUDA = -2,     /* a CUDA runtime call failed (message has the CUDA error string)  */
  B200_ERR_NO_DEVICE = -3,/* no usable sm_100 device / driver                               */
  B200_ERR_NCCL = -4,     /* NCCL missing or an NCCL call failed                            */
  B200_ERR_NOMEM = -5
} b200_status;

/* One API field of a stencil call — the C form of the (buffer, origin) pairs of the reference's
 * run_computation (backend/gtc_common.py:144-168).  The launcher borrows `data` for the call. */
typedef struct b200_field {
  void* data;                      /* device address of array element [0, 0, 0, …]; NULL = unused argument */
  int64_t strides[B200_MAX_DIMS];  /* ELEMENT strides of I, J, K, d0..d3 (0 for an axis the field lacks)  */
  int32_t origin[3];               /* array index of domain point (0,0,0) along I, J, K (0 if axis lacks) */
  int32_t shape[3];                /* array extents along I, J, K (1 if the field lacks the axis)          */
} b200_field_t;

typedef struct b200_stencil b200_stencil_t; /* a loaded stencil: device code + launch plan + scratch */
typedef struct b200_comm b200_comm_t;       /* an NCCL communicator for halo exchanges               */

/* ---- library ------------------------------------------------------------------------------ */
int b200_abi_version(void);
/* sizeof(b200_field_t) as compiled into the library (binding self-check) */
size_t b200_sizeof_field(void);
const char* b200_last_error(void);
/* Number of CUDA devices; sets *sm_major/*sm_minor of `device` when non-NULL. */
int b200_device_info(int device, int* n_devices, int* sm_major, int* sm_minor, int* n_sms);

/* ---- stencil lifecycle (replaces: importing the per-stencil pybind11 extension,
 *      reference backend/gtc_common.py:215-261 + backend/pyext_builder.py) ---------------------- */
/* `image` is an sm_100a cubin produced by the b200 code generator; `plan_text` its launch plan
 * (gt4py_b200.codegen.plan_to_text).  Both are copied. */
int b200_stencil_load(const void* image, size_t image_size, const char* plan_text, b200_stencil_t** out);
int b200_stencil_unload(b200_stencil_t* st);
int b200_stencil_num_fields(const b200_stencil_t* st);   /* API fields expected by run()            */
size_t b200_stencil_scalars_size(const b200_stencil_t* st);
int b200_stencil_num_kernels(const b200_stencil_t* st);
/* name of kernel `index` (for profiling tools); NULL when out of range */
const char* b200_stencil_kernel_name(const b200_stencil_t* st, int index);

/* ---- the hot path: one stencil application (replaces run_computation → gridtools run) -------- */
/* fields    : `nfields` API fields in stencil-signature order
 * scalars   : scalar parameters packed by the layout in the plan (`scalars_size` bytes)
 * domain    : compute domain (nI, nJ, nK)
 * subbox    : NULL for the whole domain, else {i_lo, i_hi, j_lo, j_hi} in domain coordinates — the
 *             horizontal part of the domain this call computes (used to split interior / boundary
 *             when overlapping a halo exchange)
 * stream    : cudaStream_t; kernels are enqueued, the call does not synchronise
 * returns the number of kernel launches enqueued (>= 0) or a negative status */
int b200_stencil_run(b200_stencil_t* st, const b200_field_t* fields, int nfields, const void* scalars,
                     size_t scalars_size, const int32_t domain[3], const int32_t subbox[4], void* stream);

/* The same call for a rank of a J-slab decomposition whose halo rows are written by the neighbours' b200_halo_push:
 * kernels generated with the `halo_wait` option make the tiles that read halo rows wait (on the device) until
 * *flag_lo / *flag_hi >= epoch before their first load.  A NULL flag = no neighbour on that side. */
int b200_stencil_run_halo(b200_stencil_t* st, const b200_field_t* fields, int nfields, const void* scalars,
                          size_t scalars_size, const int32_t domain[3], const int32_t subbox[4], uint64_t* flag_lo,
                          uint64_t* flag_hi, uint64_t epoch, void* stream);

/* ---- streams / events (device-side timing for exec_info and bench) --------------------------- */
int b200_stream_create(void** stream);
/* non-blocking stream with the highest (high != 0) or lowest scheduling priority of the device: halo exchange and
 * boundary strips run on high-priority streams so that their small kernels get SM slots as soon as CTAs of the
 * concurrently running interior kernel retire */
int b200_stream_create_priority(void** stream, int high);
int b200_stream_destroy(void* stream);
int b200_stream_synchronize(void* stream);
int b200_event_create(void** event);
int b200_event_destroy(void* event);
int b200_event_record(void* event, void* stream);
int b200_stream_wait_event(void* stream, void* event);
int b200_event_elapsed_ms(void* start, void* stop, float* ms); /* synchronises on `stop` */

/* ---- stencil sequences as CUDA graphs (the caller side of the hot path: a model time step is a
 *      fixed sequence of stencil calls, reference examples/cartesian/demo_burgers.ipynb cell 12; the
 *      reference pays its Python call path, stencil_object.py:296-643, on every one of them) -------
 * Everything enqueued on `stream` between begin and end — b200_stencil_run launches, halo exchanges,
 * work on streams forked/joined with b200_event_record + b200_stream_wait_event — is captured instead
 * of executed; kernel arguments are frozen at capture time.  Scratch for temporaries is owned per (stencil,
 * stream): a capture allocates what it needs, and a buffer referenced by a captured graph stays alive until
 * b200_stencil_unload.  b200_graph_launch replays the whole sequence with one driver call. */
typedef struct b200_graph b200_graph_t;
int b200_graph_begin(void* stream);
int b200_graph_end(void* stream, b200_graph_t** out);
int b200_graph_num_nodes(const b200_graph_t* graph);
int b200_graph_launch(b200_graph_t* graph, void* stream);
int b200_graph_destroy(b200_graph_t* graph);

/* ---- multi-GPU halo exchange over NCCL (an addition: the reference has no distributed path) --- */
#define B200_NCCL_UNIQUE_ID_BYTES 128
int b200_comm_unique_id(void* id_out /* B200_NCCL_UNIQUE_ID_BYTES */);
int b200_comm_init(b200_comm_t** out, const void* unique_id, int n_ranks, int rank);
int b200_comm_destroy(b200_comm_t* comm);
/* One J-slab halo exchange for one field, enqueued on `stream` inside a single NCCL group:
 * send `count` bytes at send_lo to rank-1 / send_hi to rank+1, receive into recv_lo / recv_hi.
 * A negative peer (domain edge) skips that side. */
typedef struct b200_halo {
  const void* send_lo; void* recv_lo;   /* exchanged with `peer_lo` */
  const void* send_hi; void* recv_hi;   /* exchanged with `peer_hi` */
  size_t bytes;
} b200_halo_t;
int b200_halo_exchange(b200_comm_t* comm, const b200_halo_t* halos, int n_halos, int peer_lo, int peer_hi, void* stream);
/* Halo exchange over PEER MEMORY (NVLink / NVSwitch, no NCCL call, no staging): copy `nboxes` boxes of rows from this
 * rank's memory into the neighbours' memory (`dst` = address of the neighbour's halo rows mapped into this process,
 * e.g. torch symmetric memory), then store `epoch` to each of `flags` (also in the neighbours' memory) with system-scope
 * release semantics.  The neighbour consumes the rows with b200_stencil_run_halo.  Everything is enqueued on `stream`. */
typedef struct b200_push {
  const void* src; void* dst;
  size_t row_bytes, rows, levels;                 /* per box: levels x rows rows of row_bytes bytes (4-byte multiples) */
  size_t src_row_pitch, src_level_pitch;          /* bytes */
  size_t dst_row_pitch, dst_level_pitch;
} b200_push_t;
int b200_halo_push(const b200_push_t* boxes, int nboxes, uint64_t* const* flags, int nflags, uint64_t epoch, void* stream);
/* Consumer side for stencils whose kernels were not generated with `halo_wait`: enqueue a one-thread kernel on `stream`
 * that returns once *flag_lo / *flag_hi (this rank's memory, NULL = no neighbour) have reached `epoch`. */
int b200_halo_wait(const uint64_t* flag_lo, const uint64_t* flag_hi, uint64_t epoch, void* stream);
/* Strided slab <-> contiguous staging buffer copy kernels (J-halo slabs of a (2,1,0)-layout field
 * are nK separate chunks): rows × row_bytes, source/destination pitch in bytes. */
int b200_pack_2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows, void* stream);


/* ---- data movement of the storage / host-call path (replaces the cupy copies of the reference's storage front-end,
 *      storage/cartesian/interface.py:264-327, storage/cartesian/utils.py:237-279) ------------------------------------ */
/* Copy a box of `levels` x `rows` rows of `row_bytes` bytes between two pitched buffers (host or device, any direction)
 * with ONE strided DMA (cudaMemcpy3DAsync): only the compute domain of an output travels back to the host, its halo
 * in the caller's array stays untouched.  *_pitch = bytes between rows, *_level_rows = rows between levels. */
int b200_copy_box(void* dst, size_t dst_pitch, size_t dst_level_rows, const void* src, size_t src_pitch,
                  size_t src_level_rows, size_t row_bytes, size_t rows, size_t levels, void* stream);
/* dst[i,j,k] = src[i,j,k] for a 3-D array of `itemsize`-byte elements given ELEMENT strides on both sides (device
 * memory): the upload of a C-ordered host array into the pitched I-unit-stride storage layout.  Tiled through shared
 * memory when the unit-stride axes differ (coalesced on both sides). */
int b200_relayout(void* dst, const void* src, int itemsize, const int32_t shape[3], const int64_t dst_strides[3],
                  const int64_t src_strides[3], void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GT4PY_B200_H */
