// Translation unit of k_dg_stage_split (dg2d_split.cuh): instantiations + launcher, kept apart from dg2d.cu so that the
// production stage kernel of the 2D DG path (2d/benchmark_2d_dg.f90:1137-1479 fused with :683-707 and 2d/limiters.f90:478-654)
// rebuilds in seconds.
#include <cstdlib>
#include "dg2d_common.cuh"
#include "dg2d_fast.cuh"
#include "dg2d_tma.cuh"
#include "dg2d_split.cuh"

namespace wb { namespace dg {

namespace {
constexpr int MAXDEV = 64;

template <int M, bool ANYFLUX, bool SRC, bool OUT2>
int launch1(const CUtensorMap* map, const double* in, const StageCoef& C, double* out, const double* gx, const double* gy,
            const unsigned char* fz, const DgGrid& g, const DgPhys& P, const FastBasis& B, const DgCtrl* ctrl, int onp, int rows,
            int row_begin, int row_end, cudaStream_t stream) {
  auto kern = k_dg_stage_split<M, ANYFLUX, SRC, OUT2>;
  // function attributes are per device: configure once per (kernel, device)
  static bool configured[MAXDEV] = {};
  int dev = 0;
  WB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAXDEV || !configured[dev]) {
    WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SplitLayout<M>::template bytes<SRC>()));
    if (dev >= 0 && dev < MAXDEV) configured[dev] = true;
  }
  const int nrows = row_end - row_begin;
  if (nrows <= 0) return WB_OK;
  dim3 b(128), gr((unsigned)(g.nx / 32), (unsigned)((nrows + rows - 1) / rows));
  static const int variant = getenv("WB_DG2D_SCHED") ? atoi(getenv("WB_DG2D_SCHED")) : 0;      // phase-B schedule (development knob)
  kern<<<gr, b, SplitLayout<M>::template bytes<SRC>(), stream>>>(*map, in, C, out, gx, gy, fz, g, P, B, ctrl, onp, rows | (variant << 16), row_begin,
                                                                 row_end);
  WB_LAUNCH_CHECK();
  return WB_OK;
}

template <int M>
int launch_m(const CUtensorMap* map, const double* in, const StageCoef& C, double* out, const double* gx, const double* gy,
             const unsigned char* fz, const DgGrid& g, const DgPhys& P, const FastBasis& B, const DgCtrl* ctrl, int onp, int rows,
             int row_begin, int row_end, cudaStream_t stream) {
  const bool any = P.flux_id >= 2, src = P.source != 1;
#define WB_ARGS map, in, C, out, gx, gy, fz, g, P, B, ctrl, onp, rows, row_begin, row_end, stream
  if (C.out2) {
    if (any) return src ? launch1<M, true, true, true>(WB_ARGS) : launch1<M, true, false, true>(WB_ARGS);
    return src ? launch1<M, false, true, true>(WB_ARGS) : launch1<M, false, false, true>(WB_ARGS);
  }
  if (any) return src ? launch1<M, true, true, false>(WB_ARGS) : launch1<M, true, false, false>(WB_ARGS);
  return src ? launch1<M, false, true, false>(WB_ARGS) : launch1<M, false, false, false>(WB_ARGS);
#undef WB_ARGS
}
}  // namespace

int launch_stage_split(const CUtensorMap* map, const double* in, const StageCoef& C, double* out, const double* gx,
                       const double* gy, const unsigned char* fz, const DgGrid& g, const DgPhys& P, const FastBasis& B,
                       const DgCtrl* ctrl, int onp, int rows, int row_begin, int row_end, cudaStream_t stream) {
  if (g.nx % 32 != 0) { set_error("k_dg_stage_split needs nx %% 32 == 0 (got %d)", g.nx); return WB_ERR_ARG; }
  switch (g.m) {
    case 1: return launch_m<1>(map, in, C, out, gx, gy, fz, g, P, B, ctrl, onp, rows, row_begin, row_end, stream);
    case 2: return launch_m<2>(map, in, C, out, gx, gy, fz, g, P, B, ctrl, onp, rows, row_begin, row_end, stream);
    case 3: return launch_m<3>(map, in, C, out, gx, gy, fz, g, P, B, ctrl, onp, rows, row_begin, row_end, stream);
    default: return launch_m<4>(map, in, C, out, gx, gy, fz, g, P, B, ctrl, onp, rows, row_begin, row_end, stream);
  }
}

}}  // namespace wb::dg
