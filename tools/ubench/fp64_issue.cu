// Micro-benchmark (development aid, not product): FP64 pipe of one B200 SM sub-partition.
//   lat   : dependent DFMA chain, 1 warp                      -> cycles per DFMA = latency
//   tput  : NCH independent chains, 1 warp                    -> cycles per DFMA at ILP = NCH
//   warps : 1..8 warps per SMSP, ILP 4
//   mix   : ILP-8 DFMA stream with K independent integer ops between DFMAs -> does INT issue ride along for free?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH, int NINT>
__global__ void k(double* out, long long* cyc, int iters, double a, double b, int ia) {
  double x[NCH];
  int y[4] = {ia, ia + 1, ia + 2, ia + 3};
#pragma unroll
  for (int c = 0; c < NCH; ++c) x[c] = a + c + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        x[c] = fma(x[c], b, a);
#pragma unroll
        for (int n = 0; n < NINT; ++n) y[n & 3] = y[n & 3] * 3 + ia;      // IMAD, independent of the DFMAs
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < NCH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + y[0] + y[1] + y[2] + y[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NCH, int NINT>
void run(const char* name, int warps_per_smsp) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, sizeof(double) * 4096); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<NCH, NINT><<<1, 128 * warps_per_smsp>>>(out, cyc, iters, 1.0, 0.999, 1);
  k<NCH, NINT><<<1, 128 * warps_per_smsp>>>(out, cyc, iters, 1.0, 0.999, 1);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_dfma_warp = (double)h / (iters * 8.0 * NCH);
  printf("%-6s ILP %d, %d int/DFMA, %d warps/SMSP: %.2f cycles per DFMA per warp, %.2f cycles per DFMA per SMSP\n", name, NCH, NINT,
         warps_per_smsp, per_dfma_warp, per_dfma_warp / warps_per_smsp);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1, 0>("lat", 1);
  run<2, 0>("tput", 1); run<4, 0>("tput", 1); run<8, 0>("tput", 1);
  run<1, 0>("warps", 2); run<1, 0>("warps", 4); run<1, 0>("warps", 8);
  run<2, 0>("warps", 4); run<4, 0>("warps", 4); run<4, 0>("warps", 2);
  run<8, 1>("mix", 1); run<8, 2>("mix", 1); run<8, 4>("mix", 1);
  run<4, 1>("mix", 4); run<4, 2>("mix", 4); run<4, 4>("mix", 4);
  run<1, 2>("mix", 4); run<1, 4>("mix", 4);
  return 0;
}
