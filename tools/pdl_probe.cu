// Dependent-launch gap inside a CUDA graph, with and without programmatic dependent launch (griddepcontrol).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pdl_probe tools/pdl_probe.cu && build/pdl_probe
#include <cstdio>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void step_kernel(const double *__restrict__ in, double *__restrict__ out, int n, int work) {
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double v = in[i];
    for (int k = 0; k < work; k++) v = v * 1.0000001 + 1e-9;
    out[i] = v;
  }
}

template <bool PDL>
static void launch(const double *in, double *out, int n, int work, int ctas, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(256);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = PDL ? 1 : 0;
  cudaLaunchKernelEx(&cfg, step_kernel<PDL>, in, out, n, work);
}

template <bool PDL>
static float run(int ctas, int work, bool graph) {
  const int n = ctas * 256, L = 200;
  double *a, *b;
  cudaMalloc(&a, n * 8);
  cudaMalloc(&b, n * 8);
  cudaMemset(a, 0, n * 8);
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaGraphExec_t ge = nullptr;
  if (graph) {
    cudaGraph_t g;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < L; i++) launch<PDL>(i & 1 ? b : a, i & 1 ? a : b, n, work, ctas, s);
    cudaStreamEndCapture(s, &g);
    if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1; }
  }
  float best = 1e9;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, s);
    if (graph) cudaGraphLaunch(ge, s);
    else for (int i = 0; i < L; i++) launch<PDL>(i & 1 ? b : a, i & 1 ? a : b, n, work, ctas, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("error: %s\n", cudaGetErrorString(err));
  // correctness: 200 steps of the recurrence from 0
  double h, ref = 0;
  cudaMemcpy(&h, a, 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < L * 5; i++) for (int k = 0; k < work; k++) ref = ref * 1.0000001 + 1e-9;
  printf("  value %.12e ref %.12e %s\n", h, ref, h == ref ? "ok" : "MISMATCH");
  cudaFree(a);
  cudaFree(b);
  return best * 1000.f / L;
}

int main() {
  for (int graph = 0; graph < 2; graph++)
    for (int ctas : {1, 148, 296, 1184})
      for (int work : {1, 2000, 20000}) {
        float t0 = run<false>(ctas, work, graph);
        float t1 = run<true>(ctas, work, graph);
        printf("%s ctas %4d work %5d: plain %.2f us/launch, pdl %.2f us/launch\n", graph ? "graph " : "stream", ctas, work, t0, t1);
      }
  return 0;
}
