// Shared host/device helpers for the wbeuler C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include "../../include/wbeuler.h"

namespace wb {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define WB_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      wb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                    cudaGetErrorString(_e));                                               \
      return WB_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define WB_CHECK(expr)                  \
  do {                                  \
    int _s = (expr);                    \
    if (_s != WB_OK) return _s;         \
  } while (0)

#define WB_REQUIRE(cond, ...)           \
  do {                                  \
    if (!(cond)) {                      \
      wb::set_error(__VA_ARGS__);       \
      return WB_ERR_ARG;                \
    }                                   \
  } while (0)

#define WB_LAUNCH_CHECK()               \
  do {                                  \
    wb::g_launches.fetch_add(1);        \
    WB_CUDA(cudaGetLastError());        \
  } while (0)

// Selects the device and verifies it is a Blackwell sm_100 part.  No CPU fallback exists.
int select_device(int device_ordinal, int* chosen);

// ---------------------------------------------------------------- NCCL (dlopen'ed, no link dependency)
struct Nccl;                       // opaque communicator wrapper
int nccl_get_unique_id(void* id128);
int nccl_comm_create(Nccl** c, const void* id128, int rank, int nranks);
void nccl_comm_destroy(Nccl* c);
// grouped exchange of ghost rows with the slab neighbours: send `count` doubles from send_lo to
// rank-1 / send_hi to rank+1 and receive into recv_lo / recv_hi (null pointers skip a side)
int nccl_halo_exchange(Nccl* c, int rank, int nranks, const double* send_lo, double* recv_lo,
                       const double* send_hi, double* recv_hi, size_t count, cudaStream_t s);
// generalised: lists of (ptr,count) segments per side, all inside one ncclGroup
struct HaloSeg { const double* send; double* recv; size_t count; };
int nccl_halo_exchange_multi(Nccl* c, int lo_peer, int hi_peer, const HaloSeg* lo, int nlo,
                             const HaloSeg* hi, int nhi, cudaStream_t s);
int nccl_ring_exchange(Nccl* c, int lo_peer, int hi_peer, const double* send_lo, const double* send_hi, double* recv_lo,
                       double* recv_hi, size_t count, cudaStream_t s);
int nccl_allreduce_max_u64(Nccl* c, unsigned long long* buf, size_t count, cudaStream_t s);
int nccl_allgather_bytes(Nccl* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
int nccl_allreduce_min_f64(Nccl* c, double* buf, size_t count, cudaStream_t s);
int nccl_allreduce_max_f64(Nccl* c, double* buf, size_t count, cudaStream_t s);
int nccl_bcast_f64(Nccl* c, double* buf, size_t count, int root, cudaStream_t s);

// ---------------------------------------------------------------- peer-memory ghost rows: the flag protocol
// A slab's boundary-row launch stores its rows straight into the neighbours' ghost rows (buffers mapped with CUDA IPC);
// what is left of the exchange is one flag word per direction.  peer_signal (same stream, after that launch: its stores
// are complete) writes the exchange number into the neighbours' flag words (null = no neighbour on that side);
// peer_wait spins on the rank's own words [0] (written by the neighbour below) and [1] (above) until both have reached it,
// gives up after ~60 s and records the exchange number in word [2] (checked by the *_sync entries).
int peer_signal(unsigned long long* lo_flag, unsigned long long* hi_flag, unsigned long long seq, cudaStream_t s);
int peer_wait(unsigned long long* flags, int has_lo, int has_hi, unsigned long long seq, cudaStream_t s);

// ---------------------------------------------------------------- output path (output_file of the reference drivers)
// A table of `nrows` x `ncols` doubles that a packing kernel has written to `dbuf` (device memory owned by the job from
// here on) is copied to the host on a private stream once `ready` has fired and written to `path` in the reference's
// format '(7(1PE12.5,1X))' by a host thread: the solver's stream is never blocked.  output_wait joins and reports.
struct OutputJob;
int output_start(OutputJob** job, int dev, cudaStream_t producer, double* dbuf, size_t nrows, int ncols, const char* path);
int output_wait(OutputJob** job);      // joins (no-op on nullptr); WB_OK or the error the thread met
// one number in Fortran's 1PE12.5 (12 characters + NUL): d.dddddE+ee, or d.ddddd+eee when the exponent needs three digits
void format_1pe12_5(double v, char out[16]);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// max-reduction of non-negative doubles through their (order-preserving) bit patterns
__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* addr, double v) {
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

}  // namespace wb
