// Error plumbing, device selection and the dlopen'ed NCCL shim shared by all solvers.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>   // types/enums only; the symbols are resolved at run time with dlsym
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <cmath>

namespace wb {

static thread_local char t_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int select_device(int device_ordinal, int* chosen) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    (void)cudaGetLastError();
    return WB_ERR_CUDA;
  }
  int dev = device_ordinal;
  if (dev < 0) WB_CUDA(cudaGetDevice(&dev));
  WB_REQUIRE(dev < n, "device ordinal %d out of range (%d devices)", dev, n);
  cudaDeviceProp prop;
  WB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev, prop.name,
              prop.major, prop.minor);
    return WB_ERR_CUDA;
  }
  WB_CUDA(cudaSetDevice(dev));
  if (chosen) *chosen = dev;
  return WB_OK;
}

// ------------------------------------------------------------------------------------------ NCCL
struct NcclApi {
  void* dl = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static bool g_nccl_ok = false;

static void nccl_load() {
  // If the host program (e.g. torch) already mapped a libnccl.so.2 this returns that instance.
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.dl = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    if (g_nccl.dl) break;
  }
  if (!g_nccl.dl) return;
#define WB_SYM(field, name)                                              \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.dl, name);                     \
  if (!g_nccl.field) return;
  WB_SYM(GetUniqueId, "ncclGetUniqueId")
  WB_SYM(CommInitRank, "ncclCommInitRank")
  WB_SYM(CommDestroy, "ncclCommDestroy")
  WB_SYM(GroupStart, "ncclGroupStart")
  WB_SYM(GroupEnd, "ncclGroupEnd")
  WB_SYM(Send, "ncclSend")
  WB_SYM(Recv, "ncclRecv")
  WB_SYM(AllReduce, "ncclAllReduce")
  WB_SYM(Broadcast, "ncclBroadcast")
  WB_SYM(AllGather, "ncclAllGather")
  WB_SYM(GetErrorString, "ncclGetErrorString")
#undef WB_SYM
  g_nccl_ok = true;
}

static int nccl_ready() {
  std::call_once(g_nccl_once, nccl_load);
  if (!g_nccl_ok) {
    set_error("libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbol");
    return WB_ERR_NCCL;
  }
  return WB_OK;
}

#define WB_NCCL(expr)                                                                       \
  do {                                                                                      \
    ncclResult_t _r = (expr);                                                               \
    if (_r != ncclSuccess) {                                                                \
      set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r));  \
      return WB_ERR_NCCL;                                                                   \
    }                                                                                       \
  } while (0)

struct Nccl {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};

int nccl_get_unique_id(void* id128) {
  WB_CHECK(nccl_ready());
  static_assert(sizeof(ncclUniqueId) == WB_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  WB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return WB_OK;
}

int nccl_comm_create(Nccl** c, const void* id128, int rank, int nranks) {
  WB_CHECK(nccl_ready());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  Nccl* n = new Nccl;
  n->rank = rank;
  n->nranks = nranks;
  ncclResult_t r = g_nccl.CommInitRank(&n->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    delete n;
    return WB_ERR_NCCL;
  }
  *c = n;
  return WB_OK;
}

void nccl_comm_destroy(Nccl* c) {
  if (!c) return;
  if (c->comm && g_nccl_ok) g_nccl.CommDestroy(c->comm);
  delete c;
}

// Inside an ncclGroupStart/End region an error must not return before the group is closed: every later NCCL call of the
// thread would be queued into the open group and never issued.  `r` keeps the first failure, GroupEnd always runs.
#define WB_NCCL_IN_GROUP(r, expr)                       \
  do {                                                  \
    if ((r) == ncclSuccess) (r) = (expr);               \
  } while (0)

static int nccl_group_result(ncclResult_t r, ncclResult_t end, const char* where) {
  if (r == ncclSuccess) r = end;
  if (r != ncclSuccess) {
    set_error("NCCL error in %s: %s", where, g_nccl.GetErrorString(r));
    return WB_ERR_NCCL;
  }
  return WB_OK;
}

int nccl_halo_exchange_multi(Nccl* c, int lo_peer, int hi_peer, const HaloSeg* lo, int nlo,
                             const HaloSeg* hi, int nhi, cudaStream_t s) {
  if (lo_peer < 0 && hi_peer < 0) return WB_OK;
  if (!c || !c->comm) { set_error("NCCL communicator missing"); return WB_ERR_STATE; }
  WB_NCCL(g_nccl.GroupStart());
  ncclResult_t r = ncclSuccess;
  if (lo_peer >= 0)
    for (int k = 0; k < nlo; ++k) {
      WB_NCCL_IN_GROUP(r, g_nccl.Send(lo[k].send, lo[k].count, ncclDouble, lo_peer, c->comm, s));
      WB_NCCL_IN_GROUP(r, g_nccl.Recv(lo[k].recv, lo[k].count, ncclDouble, lo_peer, c->comm, s));
    }
  if (hi_peer >= 0)
    for (int k = 0; k < nhi; ++k) {
      WB_NCCL_IN_GROUP(r, g_nccl.Send(hi[k].send, hi[k].count, ncclDouble, hi_peer, c->comm, s));
      WB_NCCL_IN_GROUP(r, g_nccl.Recv(hi[k].recv, hi[k].count, ncclDouble, hi_peer, c->comm, s));
    }
  return nccl_group_result(r, g_nccl.GroupEnd(), "nccl_halo_exchange_multi");
}

// one contiguous message per side; the order (send lo, send hi, receive hi, receive lo) pairs the messages correctly
// also when both neighbours are the same rank (periodic ring of two)
int nccl_ring_exchange(Nccl* c, int lo_peer, int hi_peer, const double* send_lo, const double* send_hi, double* recv_lo,
                       double* recv_hi, size_t count, cudaStream_t s) {
  if (lo_peer < 0 && hi_peer < 0) return WB_OK;
  if (!c || !c->comm) { set_error("NCCL communicator missing"); return WB_ERR_STATE; }
  WB_NCCL(g_nccl.GroupStart());
  ncclResult_t r = ncclSuccess;
  if (lo_peer >= 0) WB_NCCL_IN_GROUP(r, g_nccl.Send(send_lo, count, ncclDouble, lo_peer, c->comm, s));
  if (hi_peer >= 0) WB_NCCL_IN_GROUP(r, g_nccl.Send(send_hi, count, ncclDouble, hi_peer, c->comm, s));
  if (hi_peer >= 0) WB_NCCL_IN_GROUP(r, g_nccl.Recv(recv_hi, count, ncclDouble, hi_peer, c->comm, s));
  if (lo_peer >= 0) WB_NCCL_IN_GROUP(r, g_nccl.Recv(recv_lo, count, ncclDouble, lo_peer, c->comm, s));
  return nccl_group_result(r, g_nccl.GroupEnd(), "nccl_ring_exchange");
}

int nccl_halo_exchange(Nccl* c, int rank, int nranks, const double* send_lo, double* recv_lo,
                       const double* send_hi, double* recv_hi, size_t count, cudaStream_t s) {
  HaloSeg lo{send_lo, recv_lo, count}, hi{send_hi, recv_hi, count};
  int lo_peer = (rank > 0 && send_lo) ? rank - 1 : -1;
  int hi_peer = (rank < nranks - 1 && send_hi) ? rank + 1 : -1;
  return nccl_halo_exchange_multi(c, lo_peer, hi_peer, &lo, 1, &hi, 1, s);
}

int nccl_allreduce_max_u64(Nccl* c, unsigned long long* buf, size_t count, cudaStream_t s) {
  if (!c || !c->comm) { set_error("NCCL communicator missing"); return WB_ERR_STATE; }
  WB_NCCL(g_nccl.AllReduce(buf, buf, count, ncclUint64, ncclMax, c->comm, s));
  return WB_OK;
}
int nccl_allgather_bytes(Nccl* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) {
  if (!c || !c->comm) { set_error("NCCL communicator missing"); return WB_ERR_STATE; }
  WB_NCCL(g_nccl.AllGather(send, recv, bytes_per_rank, ncclChar, c->comm, s));
  return WB_OK;
}
int nccl_allreduce_min_f64(Nccl* c, double* buf, size_t count, cudaStream_t s) {
  if (!c || !c->comm) { set_error("NCCL communicator missing"); return WB_ERR_STATE; }
  WB_NCCL(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclMin, c->comm, s));
  return WB_OK;
}
int nccl_allreduce_max_f64(Nccl* c, double* buf, size_t count, cudaStream_t s) {
  if (!c || !c->comm) { set_error("NCCL communicator missing"); return WB_ERR_STATE; }
  WB_NCCL(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclMax, c->comm, s));
  return WB_OK;
}
int nccl_bcast_f64(Nccl* c, double* buf, size_t count, int root, cudaStream_t s) {
  if (!c || !c->comm) { set_error("NCCL communicator missing"); return WB_ERR_STATE; }
  WB_NCCL(g_nccl.Broadcast(buf, buf, count, ncclDouble, root, c->comm, s));
  return WB_OK;
}

// ---------------------------------------------------------------- output path
void format_1pe12_5(double v, char out[16]) {
  if (std::isnan(v)) { snprintf(out, 16, "%12s", "NaN"); return; }
  if (std::isinf(v)) { snprintf(out, 16, "%12s", v > 0 ? "Infinity" : "-Infinity"); return; }
  char tmp[32];
  snprintf(tmp, sizeof(tmp), "%.5E", v);            // [-]d.dddddE[+-]ee[e]
  char* e = strchr(tmp, 'E');
  const int ex = atoi(e + 1);
  if (ex >= 100 || ex <= -100) {                    // Fortran drops the letter: d.ddddd+eee
    *e = 0;
    char t2[64];
    snprintf(t2, sizeof(t2), "%s%c%03d", tmp, ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
    snprintf(out, 16, "%12.12s", t2);
  } else {
    snprintf(out, 16, "%12.12s", tmp);
  }
}

struct OutputJob {
  std::thread th;
  int status = WB_OK;
  std::string err;
};

static void output_thread(OutputJob* job, int dev, cudaEvent_t ready, double* dbuf, size_t nrows, int ncols, std::string path) {
  auto fail = [&](const char* what, cudaError_t e) {
    job->status = WB_ERR_CUDA;
    job->err = std::string(what) + ": " + cudaGetErrorString(e);
  };
  cudaError_t e = cudaSetDevice(dev);
  cudaStream_t s = nullptr;
  double* hbuf = nullptr;
  const size_t bytes = sizeof(double) * nrows * ncols;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(s, ready, 0);
  if (e == cudaSuccess) e = cudaMallocHost(&hbuf, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hbuf, dbuf, bytes, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) fail("output_file copy", e);
  if (e == cudaSuccess) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) { job->status = WB_ERR_ARG; job->err = "cannot open " + path; }
    else {
      std::vector<char> line((size_t)ncols * 13 + 2);
      for (size_t r = 0; r < nrows; ++r) {
        char* p = line.data();
        for (int c = 0; c < ncols; ++c) {
          char num[16];
          format_1pe12_5(hbuf[r * ncols + c], num);
          memcpy(p, num, 12); p += 12;
          if (c + 1 < ncols) *p++ = ' ';            // 1X between the fields; a trailing 1X writes nothing
        }
        *p++ = '\n';
        fwrite(line.data(), 1, (size_t)(p - line.data()), f);
      }
      fclose(f);
    }
  }
  if (hbuf) cudaFreeHost(hbuf);
  if (s) cudaStreamDestroy(s);
  cudaFree(dbuf);
  cudaEventDestroy(ready);
}

int output_start(OutputJob** job, int dev, cudaStream_t producer, double* dbuf, size_t nrows, int ncols, const char* path) {
  WB_CHECK(output_wait(job));                       // one job per handle at a time
  cudaEvent_t ready;
  WB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  WB_CUDA(cudaEventRecord(ready, producer));
  OutputJob* j = new OutputJob;
  j->th = std::thread(output_thread, j, dev, ready, dbuf, nrows, ncols, std::string(path));
  *job = j;
  return WB_OK;
}

int output_wait(OutputJob** job) {
  if (!job || !*job) return WB_OK;
  OutputJob* j = *job;
  if (j->th.joinable()) j->th.join();
  const int st = j->status;
  if (st != WB_OK) set_error("%s", j->err.c_str());
  delete j;
  *job = nullptr;
  return st;
}

// ---------------------------------------------------------------- peer-memory ghost rows: the flag protocol (common.cuh)
namespace {
__global__ void k_peer_signal(unsigned long long* lo_flag, unsigned long long* hi_flag, unsigned long long seq) {
  __threadfence_system();
  if (lo_flag) *reinterpret_cast<volatile unsigned long long*>(lo_flag) = seq;
  if (hi_flag) *reinterpret_cast<volatile unsigned long long*>(hi_flag) = seq;
  __threadfence_system();
}
__global__ void k_peer_wait(unsigned long long* flags, int has_lo, int has_hi, unsigned long long seq) {
  const volatile unsigned long long* f = flags;
  const long long t0 = clock64();
  while ((has_lo && f[0] < seq) || (has_hi && f[1] < seq)) {
    if (clock64() - t0 > 120000000000LL) { flags[2] = seq; break; }     // ~60 s of SM clocks: a neighbour died
    __nanosleep(200);
  }
  __threadfence_system();
}
}  // namespace
int peer_signal(unsigned long long* lo_flag, unsigned long long* hi_flag, unsigned long long seq, cudaStream_t s) {
  k_peer_signal<<<1, 1, 0, s>>>(lo_flag, hi_flag, seq);
  WB_LAUNCH_CHECK();
  return WB_OK;
}
int peer_wait(unsigned long long* flags, int has_lo, int has_hi, unsigned long long seq, cudaStream_t s) {
  k_peer_wait<<<1, 1, 0, s>>>(flags, has_lo, has_hi, seq);
  WB_LAUNCH_CHECK();
  return WB_OK;
}

}  // namespace wb

extern "C" {
const char* wb_last_error(void) { return wb::t_err; }
const char* wb_version(void) { return "wbeuler-b200 0.1 (sm_100a, FP64)"; }
long long wb_kernel_launch_count(void) { return wb::g_launches.load(); }
void wb_format_1pe12_5(double v, char* out13) { char b[16]; wb::format_1pe12_5(v, b); memcpy(out13, b, 13); }
int wb_nccl_get_unique_id(void* id128) {
  if (!id128) { wb::set_error("null id buffer"); return WB_ERR_ARG; }
  return wb::nccl_get_unique_id(id128);
}
}
