// Micro-probes of the B200 SMSP issue model used in DESIGN.md (integrate kernel analysis).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_probe issue_probe.cu && ./issue_probe
// Each kernel runs 8 independent chains per thread, 8 warps per block, 8 blocks per SM.
//   0 DFMA  1 IMAD  2 DFMA+IMAD  3 DMUL  4 DADD  5 DSETP+FSEL(select)  6 DFMA+DADD+DMUL mix  7 FSEL(pairs)
//   8 DFMA + FSEL pairs
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(double *sink, int iters, double x, double y) {
  double a[8];
  unsigned b[8];
  const unsigned m = threadIdx.x * 2654435761u + 12345u, c = blockIdx.x + 7u;
#pragma unroll
  for (int k = 0; k < 8; k++) { a[k] = 1.0 + k + 1e-3 * threadIdx.x; b[k] = k + threadIdx.x; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (MODE == 0 || MODE == 2 || MODE == 8) a[k] = fma(a[k], x, y);
      if (MODE == 1 || MODE == 2) b[k] = b[k] * m + c;
      if (MODE == 3) a[k] = a[k] * x;
      if (MODE == 4) a[k] = a[k] + y;
      if (MODE == 5) a[k] = (a[k] > x) ? y : a[k] + 0.0 * y;  // DSETP + 2 FSEL (+ nothing else after folding)
      if (MODE == 6) { if (k % 3 == 0) a[k] = fma(a[k], x, y); else if (k % 3 == 1) a[k] = a[k] + y; else a[k] = a[k] * x; }
      if (MODE == 7 || MODE == 8) {
        const int hi = __double2hiint(a[k]), lo = __double2loint(a[k]);
        const bool p = (b[k] & 1u) != 0;
        a[k] = __hiloint2double(p ? hi : lo, p ? lo : hi);
        if (MODE == 7) b[k] += 1;
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k] + (double)b[k];
  if (s == 12345.678) sink[(blockIdx.x * blockDim.x + threadIdx.x) & 0xffff] = s;
}

template <int MODE>
float run(double *sink, int blocks, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(sink, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float t;
    cudaEventElapsedTime(&t, e0, e1);
    if (rep && t < best) best = t;
  }
  return best;
}

int main() {
  double *sink;
  cudaMalloc(&sink, 1 << 20);
  const int blocks = 148 * 8, iters = 8192;
  const double groups = (double)blocks * 8 * iters * 8.0;  // (warp, chain, iteration) groups
  const double cyc = 148 * 4 * 1.965e9;
  const char *names[] = {"DFMA", "IMAD", "DFMA+IMAD", "DMUL", "DADD", "DSETP+select", "DFMA/DADD/DMUL mix", "FSEL pair+IADD", "DFMA+FSEL pair"};
  float ms[9] = {run<0>(sink, blocks, iters), run<1>(sink, blocks, iters), run<2>(sink, blocks, iters),
                 run<3>(sink, blocks, iters), run<4>(sink, blocks, iters), run<5>(sink, blocks, iters),
                 run<6>(sink, blocks, iters), run<7>(sink, blocks, iters), run<8>(sink, blocks, iters)};
  for (int i = 0; i < 9; i++)
    printf("%-20s %.3f ms   %.3f groups/clk/SMSP  (%.2f clk per group)\n", names[i], ms[i],
           groups / (ms[i] * 1e-3) / cyc, cyc * ms[i] * 1e-3 / groups);
  return 0;
}
