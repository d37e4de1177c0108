// DMMA issue-pattern microbenchmarks (scratch, not product code)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: operands in registers; 1: LDS each step; 2: LDS + warp-uniform predicate per fragment
template <int MI, int NJ, int MODE>
__global__ void __launch_bounds__(512, 1) k(double *out, int iters, int nf) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0;
  double a[MI], b[NJ];
#pragma unroll
  for (int i = 0; i < MI; i++) a[i] = 1.0 + i + lane * 1e-3;
#pragma unroll
  for (int j = 0; j < NJ; j++) b[j] = 0.5 + j - lane * 1e-3;
  const double *pa = sm + fk * 324 + fr, *pb = sm + 4096 + fk * 404 + fr;
  for (int it = 0; it < iters; it++) {
    if (MODE >= 1) {
      int off = (it & 15) * 18 + (it & 3);
#pragma unroll
      for (int j = 0; j < NJ; j++) b[j] = pb[(it & 7) * 16 + j * 8];
#pragma unroll
      for (int i = 0; i < MI; i++) {
        if (MODE == 2 && i >= nf) continue;
        a[i] = pa[off + i * 8 + (i >> 1) * 18];
#pragma unroll
        for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MI, int NJ, int MODE>
void run(const char *name, int threads, int ctas_per_sm, double *out) {
  int iters = 4096, blocks = 148 * ctas_per_sm;
  cudaFuncSetAttribute(k<MI, NJ, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0);
    k<MI, NJ, MODE><<<blocks, threads, 65536>>>(out, iters, MI);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  double flops = (double)blocks * (threads / 32) * iters * MI * NJ * 512.0;
  printf("%-28s thr=%3d cta/sm=%d  %6.2f TFLOP/s  (%s)\n", name, threads, ctas_per_sm, flops / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double *out; cudaMalloc(&out, 148 * 4 * 512 * 8);
#define R(MI, NJ) \
  run<MI, NJ, 0>(#MI "x" #NJ " regs", 256, 1, out); run<MI, NJ, 0>(#MI "x" #NJ " regs", 512, 1, out); \
  run<MI, NJ, 1>(#MI "x" #NJ " lds", 256, 1, out); run<MI, NJ, 1>(#MI "x" #NJ " lds", 512, 1, out); run<MI, NJ, 1>(#MI "x" #NJ " lds", 256, 2, out); \
  run<MI, NJ, 2>(#MI "x" #NJ " lds+pred", 256, 1, out); run<MI, NJ, 2>(#MI "x" #NJ " lds+pred", 512, 1, out);
  R(7, 2) R(2, 9) R(4, 4) R(7, 4) R(4, 2) R(8, 2)
  return 0;
}
