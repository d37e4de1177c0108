// TMA probe (scratch): a 4-D FP64 tile box with negative start coordinates and out-of-bounds zero fill,
// completion through an mbarrier -- the staging scheme of the convolution input tiles.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned su32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, double *out, int box_doubles, int x0, int y0, int c0, int n0) {
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(box_doubles * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(su32(sm)), "l"(&tm), "r"(su32(&bar)), "r"(x0), "r"(y0), "r"(c0), "r"(n0) : "memory");
  }
  unsigned done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(su32(&bar)), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < box_doubles; i += blockDim.x) out[i] = sm[i];
}
int main() {
  const int N = 3, P = 5, H = 14, W = 14;
  const int BX = 18, BY = 18, BC = 4, BN = 2;
  std::vector<double> h((size_t)N * P * H * W);
  for (size_t i = 0; i < h.size(); i++) h[i] = 1.0 + (double)i;
  double *d, *o;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&o, BX * BY * BC * BN * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &q);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tm;
  cuuint64_t dims[4] = {W, H, P, N};
  cuuint64_t strides[3] = {W * 8ull, (cuuint64_t)H * W * 8, (cuuint64_t)P * H * W * 8};
  cuuint32_t box[4] = {BX, BY, BC, BN}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  for (int test = 0; test < 5; test++) {
    int x0 = test == 0 ? 0 : test == 1 ? -4 : test == 2 ? -1 : test == 3 ? 1 : -3, y0 = test ? -4 : 0, c0 = test ? 4 : 0, n0 = test ? 2 : 0;
    cudaMemset(o, 0xff, BX * BY * BC * BN * 8);
    probe<<<1, 128, BX * BY * BC * BN * 8>>>(tm, o, BX * BY * BC * BN, x0, y0, c0, n0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<double> g(BX * BY * BC * BN);
    cudaMemcpy(g.data(), o, g.size() * 8, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int n = 0; n < BN; n++) for (int c = 0; c < BC; c++) for (int y = 0; y < BY; y++) for (int x = 0; x < BX; x++) {
      int gn = n0 + n, gc = c0 + c, gy = y0 + y, gx = x0 + x;
      double want = (gn < N && gc < P && gy >= 0 && gy < H && gx >= 0 && gx < W) ? h[(((size_t)gn * P + gc) * H + gy) * W + gx] : 0.0;
      if (g[((n * BC + c) * BY + y) * BX + x] != want) bad++;
    }
    printf("test %d: %ld mismatches of %zu\n", test, bad, g.size());
  }
  return 0;
}
