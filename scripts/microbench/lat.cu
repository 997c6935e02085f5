// Latency microbenchmarks on one SM (B200): barrier, dependent DFMA, shared-memory load, rsqrt, L2 load.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, const double* g, int n) {
  __shared__ double sm[4096];
  const int tid = threadIdx.x;
  for (int i = tid; i < 4096; i += blockDim.x) sm[i] = (double)((i * 7 + 1) % 4096);
  __syncthreads();
  long long t0, t1;
  // barrier
  t0 = clock64();
  for (int i = 0; i < 256; ++i) __syncthreads();
  t1 = clock64();
  if (tid == 0) out[0] = (double)(t1 - t0) / 256;
  // barrier_or
  int v = 0;
  t0 = clock64();
  for (int i = 0; i < 256; ++i) v += __syncthreads_or(tid == i);
  t1 = clock64();
  if (tid == 0) out[1] = (double)(t1 - t0) / 256 + 1e-9 * v;
  // dependent DFMA chain
  double a = out[8 + (tid & 1)], b = 1.0000001;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) { a = a * b + 1e-9; a = a * b + 1e-9; a = a * b + 1e-9; a = a * b + 1e-9; }
  t1 = clock64();
  if (tid == 0) out[2] = (double)(t1 - t0) / 256;
  out[16 + tid] = a;
  // dependent shared load chain (pointer chasing)
  int idx = tid & 4095;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) idx = (int)sm[idx];
  t1 = clock64();
  if (tid == 0) out[3] = (double)(t1 - t0) / 256;
  out[16 + 1024 + tid] = idx;
  // rsqrt chain
  double r = 1.0 + 1e-3 * tid;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) r = rsqrt(r + 1.0);
  t1 = clock64();
  if (tid == 0) out[4] = (double)(t1 - t0) / 64;
  out[16 + 2048 + tid] = r;
  // dependent global (L2) load chain: g holds indices
  int gi = tid;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) gi = (int)__ldcg(g + gi);
  t1 = clock64();
  if (tid == 0) out[5] = (double)(t1 - t0) / 64;
  out[16 + 3072 + tid] = gi;
  // division chain
  double d = 1.0 + tid;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) d = 1.0 / (d + 0.5);
  t1 = clock64();
  if (tid == 0) out[6] = (double)(t1 - t0) / 64;
  out[16 + 4096 + tid] = d;
}
int main() {
  double *out, *g; const int n = 1 << 22;
  cudaMalloc(&out, (16 + 8192) * 8); cudaMalloc(&g, n * 8ull);
  double* h = new double[n];
  for (int i = 0; i < n; ++i) h[i] = (double)((i * 1031ull + 7777) % n);
  cudaMemcpy(g, h, n * 8ull, cudaMemcpyHostToDevice);
  double init[16] = {0}; init[8] = 1.0; init[9] = 1.0; cudaMemcpy(out, init, sizeof init, cudaMemcpyHostToDevice);
  for (int nt : {32, 128, 512, 1024}) {
    k<<<1, nt>>>(out, g, n); k<<<1, nt>>>(out, g, n);
    double r[8]; cudaMemcpy(r, out, sizeof r, cudaMemcpyDeviceToHost);
    printf("threads %4d: barrier %.0f cyc, barrier_or %.0f, dependent DFMA %.1f, dependent LDS %.0f, rsqrt(double) %.0f, dependent L2 load %.0f, 1/x %.0f\n", nt, r[0], r[1], r[2], r[3], r[4], r[5], r[6]);
  }
  return 0;
}
