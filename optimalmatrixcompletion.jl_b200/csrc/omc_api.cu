// omc_api.cu -- C ABI of libomc_b200.so (include/omc_b200.h) and the small kernels around the fused
// relaxation kernel: mask compaction (K6), fused objective + MSE (K9), eigensolver self-test and the
// FP64 peak probe.  sm_100a only; there is no CPU fallback anywhere in this library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <cstdio>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/omc_b200.h"
#include "omc_device.cuh"
#include "omc_relax.cuh"
#include "omc_eigsep.cuh"
#include "omc_altmin.cuh"
#include "omc_shor.cuh"
#include "omc_big_host.h"

namespace {

thread_local std::string g_err;
cudaStream_t g_stream = nullptr;
int g_device = -1;
int g_sm_count = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(OMC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                        \
  } while (0)

#define NEED_INIT()                                                                     \
  do {                                                                                  \
    if (g_device < 0) return fail(OMC_ERR_STATE, "omc_init() has not been called");     \
    cudaSetDevice(g_device); /* the host framework may have switched the thread's GPU */ \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc(&p, count * sizeof(T));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

}  // namespace

struct omc_problem {
  int n, m, k, cut_type;
  double gamma;
  int64_t nnz;
  int Lcap;        // max cuts per node
  int state_cap;   // warm-start records
  DevBuf<double> A, Mk;
  DevBuf<unsigned long long> chunks;
  DevBuf<int> rowptr, colidx, colptr, rowidx;
  DevBuf<double> pool_x, pool_vhat;
  int pool_size = 0, pool_cap = 0;
  DevBuf<double> pool_state;
  DevBuf<double> red;   // reduction scratch for objective/MSE
  DevBuf<double> Xdev;  // staging for omc_objective_mse
  omc::StateLayout SL;
  double c0;
  // row-major copies for the batched large-block engine (omc_big.cu), built on first use
  double* AMrm = nullptr;
  unsigned char* Mkrm = nullptr;
  omcbig::ShorHost* shor = nullptr;   // Shor valid-inequality structure (omc_problem_set_shor)
  ~omc_problem() {
    if (AMrm) cudaFree(AMrm);
    if (Mkrm) cudaFree(Mkrm);
    if (shor) omcbig::big_shor_destroy(shor);
  }
};

struct omc_frontier {
  omc_problem* p;
  int B, E, Lmax, rmax, grid;
  DevBuf<int> cut_ptr, cut_ids, warm, save, status, iters, queue;
  DevBuf<uint8_t> cut_dirs;
  DevBuf<double> objective, lower_bound, res, X, Y, U, T, scratch, prof;
  omc::ScratchLayout SC;
  bool has_warm = false, has_save = false, want_T = false;
  size_t smem = 0;
  int variant = 0;
  int xs_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  omcbig::BigFrontier* big = nullptr;   // engine 2: batched large-block engine (lockstep over the frontier)
  omcbig::BigTuning tune;
};

namespace omc {

// ------------------------------------------------------------------------------------------------
// K6: mask compaction.  BitMatrix chunks (column-major bit index b = i + n j, LSB first) -> dense 0/1
// Float64 mask, row-CSR and column-CSC with ascending indices.  Bit-exact integer work.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mask_bit(const unsigned long long* chunks, long long b) {
  return (int)((chunks[b >> 6] >> (b & 63)) & 1ull);
}

__global__ void mask_expand_kernel(const unsigned long long* __restrict__ chunks, double* __restrict__ Mk, long long total) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
    Mk[e] = (double)mask_bit(chunks, e);
}

// one warp per column (which=0) or row (which=1): count, or fill ascending with ballot compaction
__global__ void mask_count_kernel(const unsigned long long* __restrict__ chunks, int n, int m, int which, int* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nseg = which ? n : m, len = which ? m : n;
  if (w >= nseg) return;
  int c = 0;
  for (int base = 0; base < len; base += 32) {
    const int t = base + lane;
    int bit = 0;
    if (t < len) bit = which ? mask_bit(chunks, (long long)w + (long long)n * t) : mask_bit(chunks, (long long)t + (long long)n * w);
    c += __popc(__ballot_sync(0xffffffffu, bit));
  }
  if (lane == 0) cnt[w + 1] = c;
  if (w == 0 && lane == 0) cnt[0] = 0;
}
__global__ void scan_inclusive_kernel(int* v, int len) {  // single block, in place, len+1 entries (v[0] = 0)
  __shared__ int carry;
  __shared__ int tmp[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 1; base <= len; base += 1024) {
    const int i = base + threadIdx.x;
    int x = (i <= len) ? v[i] : 0;
    tmp[threadIdx.x] = x;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int y = (threadIdx.x >= o) ? tmp[threadIdx.x - o] : 0;
      __syncthreads();
      tmp[threadIdx.x] += y;
      __syncthreads();
    }
    if (i <= len) v[i] = tmp[threadIdx.x] + carry;
    __syncthreads();
    if (threadIdx.x == 1023) carry += tmp[1023];
    __syncthreads();
  }
}
__global__ void mask_fill_kernel(const unsigned long long* __restrict__ chunks, int n, int m, int which,
                                 const int* __restrict__ ptr, int* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nseg = which ? n : m, len = which ? m : n;
  if (w >= nseg) return;
  int pos = ptr[w];
  for (int base = 0; base < len; base += 32) {
    const int t = base + lane;
    int bit = 0;
    if (t < len) bit = which ? mask_bit(chunks, (long long)w + (long long)n * t) : mask_bit(chunks, (long long)t + (long long)n * w);
    const unsigned bal = __ballot_sync(0xffffffffu, bit);
    if (bit) idx[pos + __popc(bal & ((1u << lane) - 1u))] = t;
    pos += __popc(bal);
  }
}

// ------------------------------------------------------------------------------------------------
// K9: fused objective + MSE.  One streaming pass over X, A (Float64) and the mask bits:
//   acc[0] = sum_I (X-A)^2, acc[1] = sum_all (X-A)^2, acc[2] = sum X^2, acc[3] = |I|.
// Deterministic two-stage reduction (per-block partials, then one block).
// Algorithmic bytes: 2*8*n*m + n*m/8.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) objective_partial_kernel(const double* __restrict__ X, const double* __restrict__ A,
                                                                const unsigned long long* __restrict__ chunks,
                                                                long long total, double* __restrict__ partial) {
  __shared__ double red[32];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x * 2;
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; e < total; e += stride) {
    if (e + 1 < total) {
      const double2 x = *reinterpret_cast<const double2*>(X + e);
      const double2 a = *reinterpret_cast<const double2*>(A + e);
      const unsigned long long wv = chunks[e >> 6] >> (e & 63);  // e even: both bits are in the same word
      const double b0 = (double)(wv & 1ull), b1 = (double)((wv >> 1) & 1ull);
      const double d0 = x.x - a.x, d1 = x.y - a.y;
      a0 += b0 * d0 * d0 + b1 * d1 * d1;
      a1 += d0 * d0 + d1 * d1;
      a2 += x.x * x.x + x.y * x.y;
      a3 += b0 + b1;
    } else {
      const double x = X[e], a = A[e];
      const double b0 = (double)mask_bit(chunks, e);
      const double d0 = x - a;
      a0 += b0 * d0 * d0;
      a1 += d0 * d0;
      a2 += x * x;
      a3 += b0;
    }
  }
  a0 = block_sum(a0, red);
  a1 = block_sum(a1, red);
  a2 = block_sum(a2, red);
  a3 = block_sum(a3, red);
  if (threadIdx.x == 0) {
    partial[4 * blockIdx.x + 0] = a0;
    partial[4 * blockIdx.x + 1] = a1;
    partial[4 * blockIdx.x + 2] = a2;
    partial[4 * blockIdx.x + 3] = a3;
  }
}
__global__ void objective_final_kernel(const double* __restrict__ partial, int nblocks, double gamma, long long total,
                                       double* __restrict__ out4) {
  __shared__ double red[32];
  double a[4] = {0, 0, 0, 0};
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
    for (int q = 0; q < 4; ++q) a[q] += partial[4 * b + q];
  for (int q = 0; q < 4; ++q) a[q] = block_sum(a[q], red);
  if (threadIdx.x == 0) {
    const double sse_in = a[0], sse_all = a[1], sxx = a[2], cnt = a[3];
    out4[0] = 0.5 * sse_in + sxx / (2.0 * gamma);                                   // OMC.jl:2352-2358
    out4[1] = (cnt == 0.0) ? 0.0 : sse_in / cnt;                                    // OMC.jl:2389-2397
    out4[2] = ((double)total == cnt) ? 0.0 : (sse_all - sse_in) / ((double)total - cnt);  // OMC.jl:2380-2388
    out4[3] = sse_all / (double)total;                                              // OMC.jl:2398-2402
  }
}

// ------------------------------------------------------------------------------------------------
// Eigensolver self-test kernel: one CTA per matrix.  Cold Jacobi, then the warm path on the same
// matrix (DMMA pre-rotation + Jacobi) and the DMMA reconstruction of the PSD projection -- the same
// device functions the relaxation kernel uses.
// ------------------------------------------------------------------------------------------------
template <int NT, int KMAX>
__global__ void __launch_bounds__(NT, 1) psd_project_debug_kernel(int N, int B, const double* __restrict__ Vin,
                                                                 double* __restrict__ Pout, double* __restrict__ lamout,
                                                                 int* __restrict__ sweeps_out, double tol) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Geo g = make_geo(N);
  const int NP = g.NP, ld = g.ld, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  double* buf0 = reinterpret_cast<double*>(smem_raw);
  double* buf1 = buf0 + (size_t)NP * ld;
  double* lam = buf1 + (size_t)NP * ld;
  double* jcs = lam + NP;
  double* jsn = jcs + NP / 2;
  double* red = jsn + NP / 2;
  int* jrot = reinterpret_cast<int*>(red + 32);
  for (int mtx = blockIdx.x; mtx < B; mtx += gridDim.x) {
    const double* V = Vin + (size_t)mtx * N * N;
    int total_sweeps = 0;
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
      for (int e = tid; e < NP * NP; e += NT) {
        const int r = e / NP, c = e - r * NP;
        double v = 0.0;
        if (r < N && c < N) v = 0.5 * (V[(size_t)r * N + c] + V[(size_t)c * N + r]);
        buf0[(size_t)r * ld + c] = v;
        if (pass == 0) buf1[(size_t)r * ld + c] = (r == c) ? 1.0 : 0.0;
      }
      __syncthreads();
      if (pass == 1) {
        gemm_rows_inplace<KMAX>(buf0, buf1, NP, ld);
        gemm_cols_inplace<KMAX>(buf0, buf1, NP, ld);
      }
      const int sw = jacobi_sym(buf0, buf1, NP, ld, tol, 40, jcs, jsn, jrot, red);
      total_sweeps += (pass == 0) ? sw : 100 * sw;
    }
    for (int i = tid; i < NP; i += NT) lam[i] = buf0[(size_t)i * ld + i];
    __syncthreads();
    for (int i = tid; i < N; i += NT) lamout[(size_t)mtx * N + i] = lam[i];
    if (tid == 0) sweeps_out[mtx] = total_sweeps;
    // P = sum_{lam>0} lam q q' via DMMA on all tiles (weights max(lam,0), every index kept)
    {
      const int T = NP >> 3, KS = NP >> 2, g_ = lane >> 2, t_ = lane & 3;
      for (int tl = warp; tl < T * T; tl += NW) {
        const int rt = tl / T, ct = tl - rt * T;
        double c0 = 0.0, c1 = 0.0;
        const double* arow = buf1 + (size_t)(rt * 8 + g_) * ld;
        const double* brow = buf1 + (size_t)(ct * 8 + g_) * ld;
        for (int kk = 0; kk < KS; ++kk) {
          const int col = kk * 4 + t_;
          dmma884(c0, c1, arow[col] * fmax(lam[col], 0.0), brow[col], c0, c1);
        }
        const int rr = rt * 8 + g_, cc = ct * 8 + 2 * t_;
        if (rr < N && cc < N) Pout[(size_t)mtx * N * N + (size_t)rr * N + cc] = c0;
        if (rr < N && cc + 1 < N) Pout[(size_t)mtx * N * N + (size_t)rr * N + cc + 1] = c1;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// FP64 peak probes (roofline denominators that MEASURED_PEAKS.json does not carry)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q) dmma884(c[q][0], c[q][1], a, b, c[q][0], c[q][1]);
  }
  double s = 0.0;
  for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace omc

// ==================================================================================================
// C ABI
// ==================================================================================================
extern "C" {

const char* omc_last_error(void) { return g_err.c_str(); }

int32_t omc_init(int32_t device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(OMC_ERR_CUDA, "no CUDA device: %s (libomc_b200 has no CPU fallback)", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(OMC_ERR_ARG, "device %d out of range [0,%d)", device, count);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(OMC_ERR_UNSUPPORTED, "device %d is sm_%d%d; libomc_b200 is built for sm_100a only", device, prop.major,
                prop.minor);
  if (g_stream != nullptr && g_device != device) {  // the stream belongs to the previously selected GPU
    cudaStreamSynchronize(g_stream);
    cudaStreamDestroy(g_stream);
    g_stream = nullptr;
  }
  if (g_stream == nullptr) CU(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  g_device = device;
  g_sm_count = prop.multiProcessorCount;
  return OMC_OK;
}

int32_t omc_shutdown(void) {
  if (g_stream) {
    cudaStreamSynchronize(g_stream);
    cudaStreamDestroy(g_stream);
    g_stream = nullptr;
  }
  g_device = -1;
  return OMC_OK;
}

void* omc_stream(void) { return (void*)g_stream; }

int32_t omc_build_flags(void) {
  int32_t f = 0;
#ifdef OMC_INFEASIBILITY_CERTIFICATE
  f |= OMC_BUILD_INFEASIBILITY_CERTIFICATE;
#endif
#ifdef OMC_INFEASIBLE_BY_BOUND
  f |= OMC_BUILD_INFEASIBLE_BY_BOUND;
#endif
  return f;
}

int32_t omc_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* free_bytes) {
  NEED_INIT();
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, g_device));
  size_t fr = 0, tot = 0;
  CU(cudaMemGetInfo(&fr, &tot));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (free_bytes) *free_bytes = (int64_t)fr;
  return OMC_OK;
}

int32_t omc_problem_create(int32_t n, int32_t m, int32_t k, const double* A, const uint64_t* mask_chunks, double gamma,
                           int32_t cut_type, int32_t state_pool_capacity, omc_problem** out) {
  NEED_INIT();
  if (!out || !A || !mask_chunks) return fail(OMC_ERR_ARG, "null argument");
  if (n <= 0 || m <= 0 || k <= 0 || k > n) return fail(OMC_ERR_ARG, "bad sizes n=%d m=%d k=%d", n, m, k);
  if (n > m) return fail(OMC_ERR_ARG, "Input matrix A must have size (n, m) with n <= m (OMC.jl:249)");
  if (!(gamma > 0)) return fail(OMC_ERR_ARG, "gamma must be positive");
  if (cut_type < 0 || cut_type > 2) return fail(OMC_ERR_ARG, "cut_type must be 0 (linear), 1 (linear2) or 2 (linear3)");
  omc_problem* p = new omc_problem();
  p->n = n; p->m = m; p->k = k; p->cut_type = cut_type; p->gamma = gamma;
  // cuts per node (= depth): P^k children per split keep trees with large k shallow; n + m > 104 needs the shared memory
  p->Lcap = (n + m > 104) ? 32 : 64;
  p->state_cap = state_pool_capacity > 0 ? state_pool_capacity : 0;
  const long long total = (long long)n * m;
  const size_t nchunks = (size_t)((total + 63) / 64);
  cudaError_t e;
#define PC(call)                                                                                          \
  do {                                                                                                    \
    e = (call);                                                                                           \
    if (e != cudaSuccess) {                                                                               \
      delete p;                                                                                           \
      return fail(OMC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e));                           \
    }                                                                                                     \
  } while (0)
  PC(p->A.alloc(total));
  PC(p->Mk.alloc(total));
  PC(p->chunks.alloc(nchunks + 1));
  PC(cudaMemsetAsync(p->chunks.p, 0, (nchunks + 1) * 8, g_stream));
  PC(cudaMemcpyAsync(p->A.p, A, total * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  PC(cudaMemcpyAsync(p->chunks.p, mask_chunks, nchunks * 8, cudaMemcpyHostToDevice, g_stream));
  PC(p->rowptr.alloc(n + 1));
  PC(p->colptr.alloc(m + 1));
  {
    const int blocks = (int)((total + 255) / 256 < g_sm_count * 8 ? (total + 255) / 256 : g_sm_count * 8);
    omc::mask_expand_kernel<<<blocks, 256, 0, g_stream>>>(p->chunks.p, p->Mk.p, total);
    omc::mask_count_kernel<<<(m * 32 + 255) / 256, 256, 0, g_stream>>>(p->chunks.p, n, m, 0, p->colptr.p);
    omc::mask_count_kernel<<<(n * 32 + 255) / 256, 256, 0, g_stream>>>(p->chunks.p, n, m, 1, p->rowptr.p);
    omc::scan_inclusive_kernel<<<1, 1024, 0, g_stream>>>(p->colptr.p, m);
    omc::scan_inclusive_kernel<<<1, 1024, 0, g_stream>>>(p->rowptr.p, n);
  }
  int nnz32 = 0;
  PC(cudaMemcpyAsync(&nnz32, p->colptr.p + m, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  PC(cudaStreamSynchronize(g_stream));
  p->nnz = nnz32;
  PC(p->colidx.alloc(nnz32 > 0 ? nnz32 : 1));
  PC(p->rowidx.alloc(nnz32 > 0 ? nnz32 : 1));
  omc::mask_fill_kernel<<<(m * 32 + 255) / 256, 256, 0, g_stream>>>(p->chunks.p, n, m, 0, p->colptr.p, p->rowidx.p);
  omc::mask_fill_kernel<<<(n * 32 + 255) / 256, 256, 0, g_stream>>>(p->chunks.p, n, m, 1, p->rowptr.p, p->colidx.p);
  // c0 = 1/2 sum_I A^2 (constant of the relaxation's dual objective), computed by the objective kernel with X = 0
  PC(p->red.alloc(4 * g_sm_count * 4 + 8));
  PC(p->Xdev.alloc(total));
  PC(cudaMemsetAsync(p->Xdev.p, 0, total * sizeof(double), g_stream));
  omc::objective_partial_kernel<<<g_sm_count * 4, 256, 0, g_stream>>>(p->Xdev.p, p->A.p, p->chunks.p, total, p->red.p);
  omc::objective_final_kernel<<<1, 256, 0, g_stream>>>(p->red.p, g_sm_count * 4, gamma, total, p->red.p + 4 * g_sm_count * 4);
  double o4[4];
  PC(cudaMemcpyAsync(o4, p->red.p + 4 * g_sm_count * 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  PC(cudaStreamSynchronize(g_stream));
  PC(cudaGetLastError());
  p->c0 = o4[0];  // X = 0: objective = 1/2 sum_I A^2
  p->SL = omc::make_state_layout(n, m, k, p->Lcap);
  if (p->state_cap > 0) PC(p->pool_state.alloc((size_t)p->state_cap * p->SL.total));
  p->pool_cap = 1024;
  PC(p->pool_x.alloc((size_t)p->pool_cap * n));
  PC(p->pool_vhat.alloc((size_t)p->pool_cap * k));
#undef PC
  *out = p;
  return OMC_OK;
}

int32_t omc_problem_destroy(omc_problem* p) {
  if (!p) return OMC_OK;
  if (g_stream) cudaStreamSynchronize(g_stream);
  delete p;
  return OMC_OK;
}

int32_t omc_problem_get_csr(omc_problem* p, int32_t* rowptr, int32_t* colidx, int32_t* colptr, int32_t* rowidx,
                            int64_t* nnz) {
  NEED_INIT();
  if (!p) return fail(OMC_ERR_ARG, "null problem");
  if (nnz) *nnz = p->nnz;
  if (rowptr) CU(cudaMemcpyAsync(rowptr, p->rowptr.p, (p->n + 1) * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  if (colptr) CU(cudaMemcpyAsync(colptr, p->colptr.p, (p->m + 1) * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  if (colidx && p->nnz) CU(cudaMemcpyAsync(colidx, p->colidx.p, p->nnz * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  if (rowidx && p->nnz) CU(cudaMemcpyAsync(rowidx, p->rowidx.p, p->nnz * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  return OMC_OK;
}

int32_t omc_cutpool_add(omc_problem* p, const double* x, const double* vhat, int32_t* cut_id) {
  NEED_INIT();
  if (!p || !x || !vhat || !cut_id) return fail(OMC_ERR_ARG, "null argument");
  if (p->pool_size == p->pool_cap) {  // grow x2
    const int ncap = p->pool_cap * 2;
    DevBuf<double> nx, nv;
    CU(nx.alloc((size_t)ncap * p->n));
    CU(nv.alloc((size_t)ncap * p->k));
    CU(cudaMemcpyAsync(nx.p, p->pool_x.p, (size_t)p->pool_size * p->n * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    CU(cudaMemcpyAsync(nv.p, p->pool_vhat.p, (size_t)p->pool_size * p->k * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    std::swap(p->pool_x.p, nx.p); std::swap(p->pool_x.n, nx.n);
    std::swap(p->pool_vhat.p, nv.p); std::swap(p->pool_vhat.n, nv.n);
    p->pool_cap = ncap;
  }
  const int id = p->pool_size;
  CU(cudaMemcpyAsync(p->pool_x.p + (size_t)id * p->n, x, p->n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(p->pool_vhat.p + (size_t)id * p->k, vhat, p->k * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  p->pool_size = id + 1;
  *cut_id = id;
  return OMC_OK;
}

int32_t omc_cutpool_size(omc_problem* p, int32_t* size) {
  if (!p || !size) return fail(OMC_ERR_ARG, "null argument");
  *size = p->pool_size;
  return OMC_OK;
}

void omc_relax_default_opts(omc_relax_opts* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->eps_abs = 1e-8;
  o->eps_rel = 1e-8;
  o->max_iter = 20000;
  o->check_every = 25;
  o->adapt_every = 100;
  o->fix_linear3_right = 0;
  o->rho0 = 0.3;   /* measured on config-2 frontier nodes: ~2x fewer ADMM iterations than 0.1 (scripts/experiments/admm_params.py) */
  o->sigma = 1e-6;
  o->alpha = 1.6;
  o->cutoff = INFINITY;
  o->time_limit_s = 0.0;
  o->jacobi_tol = 1e-3;   /* cap only: the tolerance follows the ADMM residual (1e-2 x relative residual) */
  o->reortho_every = 0;
}

int32_t omc_frontier_create(omc_problem* p, int32_t B, const int32_t* node_cut_ptr, const int32_t* node_cut_ids,
                            const uint8_t* node_cut_dirs, const int32_t* warm_ids, const int32_t* save_ids,
                            omc_frontier** out) {
  return omc_frontier_create_ex(p, B, node_cut_ptr, node_cut_ids, node_cut_dirs, warm_ids, save_ids, OMC_ENGINE_AUTO, out);
}

int32_t omc_frontier_create_ex(omc_problem* p, int32_t B, const int32_t* node_cut_ptr, const int32_t* node_cut_ids,
                               const uint8_t* node_cut_dirs, const int32_t* warm_ids, const int32_t* save_ids,
                               int32_t engine, omc_frontier** out) {
  NEED_INIT();
  if (!p || !out || B <= 0 || !node_cut_ptr) return fail(OMC_ERR_ARG, "bad argument");
  if (engine < OMC_ENGINE_AUTO || engine > OMC_ENGINE_BATCHED) return fail(OMC_ERR_ARG, "engine must be 0 (auto), 1 (persistent) or 2 (batched)");
  const int engine_asked = engine;
  const int E = node_cut_ptr[B];
  if (E < 0 || node_cut_ptr[0] != 0) return fail(OMC_ERR_ARG, "node_cut_ptr must start at 0 and be non-decreasing");
  if (E > 0 && (!node_cut_ids || !node_cut_dirs)) return fail(OMC_ERR_ARG, "cut ids/dirs missing");
  int Lmax = 0;
  const int ndir = p->cut_type + 2;
  for (int b = 0; b < B; ++b) {
    const int L = node_cut_ptr[b + 1] - node_cut_ptr[b];
    if (L < 0) return fail(OMC_ERR_ARG, "node_cut_ptr not monotone at node %d", b);
    if (L > Lmax) Lmax = L;
  }
  // auto: PSD blocks beyond one SM's shared memory, or a node deeper than the persistent engine's cut capacity -> batched engine
  if (engine == OMC_ENGINE_AUTO) engine = (p->n + p->m > 104 || Lmax > p->Lcap || p->shor) ? OMC_ENGINE_BATCHED : OMC_ENGINE_PERSISTENT;
  if (engine == OMC_ENGINE_PERSISTENT && p->shor)
    return fail(OMC_ERR_UNSUPPORTED, "Shor valid inequalities are rows of the batched engine only (engine = 0 or 2)");
  (void)engine_asked;
  if (engine == OMC_ENGINE_PERSISTENT && Lmax > p->Lcap)
    return fail(OMC_ERR_UNSUPPORTED, "node with %d cuts exceeds the %d the persistent engine supports (the batched engine, engine = 2, has no such cap)", Lmax, p->Lcap);
  for (int e = 0; e < E; ++e) {
    if (node_cut_ids[e] < 0 || node_cut_ids[e] >= p->pool_size) return fail(OMC_ERR_ARG, "cut id %d out of range", node_cut_ids[e]);
    for (int j = 0; j < p->k; ++j)
      if (node_cut_dirs[(size_t)e * p->k + j] >= ndir) return fail(OMC_ERR_ARG, "direction code out of range");
  }
  for (int b = 0; b < B; ++b) {
    if (warm_ids && warm_ids[b] >= p->state_cap) return fail(OMC_ERR_ARG, "warm id out of range");
    if (save_ids && save_ids[b] >= p->state_cap) return fail(OMC_ERR_ARG, "save id out of range");
  }
  if (save_ids && p->state_cap > 0) {   // a record written by one node must not be read or written by another node of the same launch
    std::vector<char> used((size_t)p->state_cap, 0);
    for (int b = 0; b < B; ++b)
      if (save_ids[b] >= 0) {
        if (used[save_ids[b]]) return fail(OMC_ERR_ARG, "save id %d appears twice in one frontier", save_ids[b]);
        used[save_ids[b]] = 1;
      }
    if (warm_ids)
      for (int b = 0; b < B; ++b)
        if (warm_ids[b] >= 0 && used[warm_ids[b]] && save_ids[b] != warm_ids[b])
          return fail(OMC_ERR_ARG, "state record %d is read by node %d and written by another node of the same frontier", warm_ids[b], b);
  }
  if (engine == OMC_ENGINE_BATCHED) {
    if (warm_ids || save_ids) {
      for (int b = 0; b < B; ++b)
        if ((warm_ids && warm_ids[b] >= 0) || (save_ids && save_ids[b] >= 0))
          return fail(OMC_ERR_UNSUPPORTED, "the batched engine starts every node cold (warm_ids / save_ids must be -1)");
    }
    if (!p->AMrm) {
      if (omcbig::big_prepare_problem(p->n, p->m, p->A.p, p->Mk.p, &p->AMrm, &p->Mkrm, g_stream) != 0)
        return fail(OMC_ERR_CUDA, "%s", omcbig::big_last_error());
    }
    omc_frontier* f = new omc_frontier();
    f->p = p; f->B = B; f->E = E; f->Lmax = Lmax;
    omcbig::big_default_tuning(&f->tune);
    omcbig::BigProblemView pv;
    pv.n = p->n; pv.m = p->m; pv.k = p->k; pv.cut_type = p->cut_type; pv.gamma = p->gamma; pv.c0 = p->c0;
    pv.AMrm = p->AMrm; pv.Mkrm = p->Mkrm; pv.pool_x = p->pool_x.p; pv.pool_vhat = p->pool_vhat.p; pv.stream = g_stream;
    pv.sm_count = g_sm_count;
    pv.shor = p->shor;
    const int rc = omcbig::big_create(pv, B, node_cut_ptr, node_cut_ids, node_cut_dirs, &f->big);
    if (rc != 0) {
      delete f;
      return fail(rc == -4 ? OMC_ERR_UNSUPPORTED : OMC_ERR_CUDA, "%s", omcbig::big_last_error());
    }
    *out = f;
    return OMC_OK;
  }
  const omc::Geo g1 = omc::make_geo(p->n + p->m);
  const omc::Geo g2 = omc::make_geo(p->n + p->k);
  if (g1.NP > 208 || g2.NP > OMC_SMEM_NP_MAX)
    return fail(OMC_ERR_UNSUPPORTED, "n+m = %d, n+k = %d: this build supports n+m <= 208 (L2-resident working buffers above 104) and n+k <= 104",
                p->n + p->m, p->n + p->k);
  omc_frontier* f = new omc_frontier();
  f->p = p; f->B = B; f->E = E; f->Lmax = Lmax;
  f->rmax = 1 + p->Lcap * (p->k + 1);
  // the Gram staging C[L][L] and (when it fits) the Woodbury inverse use buf0
  f->variant = (g1.NP <= 32) ? 0 : (g1.NP <= 64 ? 1 : (g1.NP <= 104 ? 2 : 3));
  const int ctas_per_sm = (f->variant == 0) ? 4 : (f->variant == 1 ? 2 : 1);
  f->grid = g_sm_count * ctas_per_sm;
  if (f->grid > B) f->grid = B;
  f->SC = omc::make_scratch_layout(p->SL, f->rmax);
  // cut vectors are cached in shared memory except where the panels of the tracked projection need the room (n + m > 104)
  f->xs_cap = (f->variant == 3) ? 0 : OMC_XS_CAP;
  f->smem = (f->variant == 0) ? omc::relax_smem_bytes<8>(p->n, p->m, p->k, p->Lcap, f->rmax, f->xs_cap)
                              : omc::relax_smem_bytes<16>(p->n, p->m, p->k, p->Lcap, f->rmax, f->xs_cap);
  if (f->smem > 227 * 1024) {
    const size_t need = f->smem;
    delete f;
    return fail(OMC_ERR_UNSUPPORTED, "shared memory need %zu bytes exceeds 227 KB for n = %d, m = %d, k = %d", need, p->n, p->m, p->k);
  }
  cudaError_t e;
#define FC(call)                                                                \
  do {                                                                          \
    e = (call);                                                                 \
    if (e != cudaSuccess) {                                                     \
      delete f;                                                                 \
      return fail(OMC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e)); \
    }                                                                           \
  } while (0)
  FC(f->cut_ptr.alloc(B + 1));
  FC(f->cut_ids.alloc(E > 0 ? E : 1));
  FC(f->cut_dirs.alloc(E > 0 ? (size_t)E * p->k : 1));
  FC(f->status.alloc(B)); FC(f->iters.alloc(B)); FC(f->queue.alloc(1));
  FC(f->objective.alloc(B)); FC(f->lower_bound.alloc(B)); FC(f->res.alloc(2 * (size_t)B));
  FC(f->X.alloc((size_t)B * p->n * p->m));
  FC(f->Y.alloc((size_t)B * p->n * p->n));
  FC(f->U.alloc((size_t)B * p->n * p->k));
  FC(f->scratch.alloc((size_t)f->grid * f->SC.total));
  FC(f->prof.alloc((size_t)B * OMC_PROF_STRIDE));
  FC(cudaMemcpyAsync(f->cut_ptr.p, node_cut_ptr, (B + 1) * sizeof(int), cudaMemcpyHostToDevice, g_stream));
  if (E > 0) {
    FC(cudaMemcpyAsync(f->cut_ids.p, node_cut_ids, E * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    FC(cudaMemcpyAsync(f->cut_dirs.p, node_cut_dirs, (size_t)E * p->k, cudaMemcpyHostToDevice, g_stream));
  }
  if (warm_ids) {
    FC(f->warm.alloc(B));
    FC(cudaMemcpyAsync(f->warm.p, warm_ids, B * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    f->has_warm = true;
  }
  if (save_ids) {
    FC(f->save.alloc(B));
    FC(cudaMemcpyAsync(f->save.p, save_ids, B * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    f->has_save = true;
  }
  FC(cudaEventCreate(&f->ev0));
  FC(cudaEventCreate(&f->ev1));
  FC(cudaStreamSynchronize(g_stream));
#undef FC
  *out = f;
  return OMC_OK;
}

int32_t omc_frontier_relax(omc_frontier* f, const omc_relax_opts* opts, float* kernel_ms) {
  NEED_INIT();
  if (!f) return fail(OMC_ERR_ARG, "null frontier");
  omc_problem* p = f->p;
  omc_relax_opts o;
  if (opts) o = *opts; else omc_relax_default_opts(&o);
  if (o.check_every <= 0) o.check_every = 25;
  if (o.adapt_every > 0) o.adapt_every = ((o.adapt_every + o.check_every - 1) / o.check_every) * o.check_every;
  if (o.max_iter <= 0) return fail(OMC_ERR_ARG, "max_iter must be positive");
  if (!(o.rho0 > 0) || !(o.sigma > 0) || !(o.alpha > 0 && o.alpha < 2)) return fail(OMC_ERR_ARG, "bad rho0/sigma/alpha");
  if (f->big) {
    if (omcbig::big_relax(f->big, &o, &f->tune, kernel_ms) != 0) return fail(OMC_ERR_CUDA, "%s", omcbig::big_last_error());
    return OMC_OK;
  }
  omc::RelaxArgs a;
  memset(&a, 0, sizeof a);
  a.n = p->n; a.m = p->m; a.k = p->k; a.cut_type = p->cut_type;
  a.gamma = p->gamma; a.a = (double)p->n; a.sa = sqrt((double)p->n); a.cT = a.a / (2.0 * p->gamma); a.c0 = p->c0;
  a.A = p->A.p; a.Mk = p->Mk.p; a.pool_x = p->pool_x.p; a.pool_vhat = p->pool_vhat.p;
  a.B = f->B; a.node_cut_ptr = f->cut_ptr.p; a.node_cut_ids = f->cut_ids.p; a.node_cut_dirs = f->cut_dirs.p;
  a.warm_ids = f->has_warm ? f->warm.p : nullptr;
  a.save_ids = f->has_save ? f->save.p : nullptr;
  a.pool_state = p->pool_state.p; a.scratch = f->scratch.p; a.queue = f->queue.p;
  a.status = f->status.p; a.objective = f->objective.p; a.lower_bound = f->lower_bound.p; a.iters = f->iters.p;
  a.prof = f->prof.p;
  a.res = f->res.p; a.outX = f->X.p; a.outY = f->Y.p; a.outU = f->U.p; a.outT = f->T.p;
  a.o = o; a.Lcap = p->Lcap; a.rmax = f->rmax; a.xs_cap = f->xs_cap; a.SL = p->SL; a.SC = f->SC;
  CU(cudaMemsetAsync(f->queue.p, 0, sizeof(int), g_stream));
  CU(cudaEventRecord(f->ev0, g_stream));
  if (f->variant == 0) {
    auto kern = omc::omc_relax_kernel<128, 8, 4, 8>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem));
    kern<<<f->grid, 128, f->smem, g_stream>>>(a);
  } else if (f->variant == 1) {
    auto kern = omc::omc_relax_kernel<256, 16, 2, 16>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem));
    kern<<<f->grid, 256, f->smem, g_stream>>>(a);
  } else if (f->variant == 2) {
    auto kern = omc::omc_relax_kernel<512, 26, 1, 16>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem));
    kern<<<f->grid, 512, f->smem, g_stream>>>(a);
  } else {  // (n+m) block in L2-resident buffers: 256 threads so that a 52-deep DMMA panel fits in registers
    auto kern = omc::omc_relax_kernel<256, 52, 1, 16>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem));
    kern<<<f->grid, 256, f->smem, g_stream>>>(a);
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(f->ev1, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  if (kernel_ms) CU(cudaEventElapsedTime(kernel_ms, f->ev0, f->ev1));
  return OMC_OK;
}

int32_t omc_frontier_fetch(omc_frontier* f, int32_t* status, double* objective, double* lower_bound, int32_t* iters,
                           double* res, double* X, double* Y, double* U, double* Theta) {
  NEED_INIT();
  if (!f) return fail(OMC_ERR_ARG, "null frontier");
  omc_problem* p = f->p;
  const size_t B = f->B;
  if (Theta) return fail(OMC_ERR_UNSUPPORTED, "Theta is not materialised by this build (only tr Theta enters the objective)");
  if (f->big) {
    if (omcbig::big_fetch(f->big, status, objective, lower_bound, iters, res, X, Y, U) != 0)
      return fail(OMC_ERR_CUDA, "%s", omcbig::big_last_error());
    return OMC_OK;
  }
  if (status) CU(cudaMemcpyAsync(status, f->status.p, B * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  if (iters) CU(cudaMemcpyAsync(iters, f->iters.p, B * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  if (objective) CU(cudaMemcpyAsync(objective, f->objective.p, B * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  if (lower_bound) CU(cudaMemcpyAsync(lower_bound, f->lower_bound.p, B * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  if (res) CU(cudaMemcpyAsync(res, f->res.p, 2 * B * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  if (X) CU(cudaMemcpyAsync(X, f->X.p, B * p->n * p->m * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  if (Y) CU(cudaMemcpyAsync(Y, f->Y.p, B * p->n * p->n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  if (U) CU(cudaMemcpyAsync(U, f->U.p, B * p->n * p->k * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  return OMC_OK;
}

int32_t omc_frontier_fetch_profile(omc_frontier* f, double* prof) {
  NEED_INIT();
  if (!f || !prof) return fail(OMC_ERR_ARG, "null argument");
  if (f->big) {   // the batched engine keeps launch statistics instead of per-node cycle counters (omc_frontier_stats)
    memset(prof, 0, (size_t)f->B * OMC_PROF_STRIDE * sizeof(double));
    return OMC_OK;
  }
  CU(cudaMemcpyAsync(prof, f->prof.p, (size_t)f->B * OMC_PROF_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  return OMC_OK;
}

int32_t omc_frontier_destroy(omc_frontier* f) {
  if (!f) return OMC_OK;
  if (g_stream) cudaStreamSynchronize(g_stream);
  if (f->ev0) cudaEventDestroy(f->ev0);
  if (f->ev1) cudaEventDestroy(f->ev1);
  if (f->big) omcbig::big_destroy(f->big);
  delete f;
  return OMC_OK;
}

int32_t omc_problem_set_shor(omc_problem* p, int64_t n_minors, const int32_t* minors, int64_t n_soc, const int32_t* soc) {
  NEED_INIT();
  if (!p || n_minors < 0 || n_soc < 0 || (n_minors > 0 && !minors) || (n_soc > 0 && !soc)) return fail(OMC_ERR_ARG, "bad argument");
  if (p->shor) { omcbig::big_shor_destroy(p->shor); p->shor = nullptr; }
  if (n_minors == 0 && n_soc == 0) return OMC_OK;
  if (n_minors > ((int64_t)1 << 29)) return fail(OMC_ERR_UNSUPPORTED, "more than 2^29 Shor minors");
  const int rc = omcbig::big_shor_create(p->n, p->m, p->k, n_minors, minors, n_soc, soc, &p->shor);
  if (rc != 0) return fail(rc == -1 ? OMC_ERR_ARG : rc == -4 ? OMC_ERR_UNSUPPORTED : OMC_ERR_CUDA, "%s", omcbig::big_last_error());
  return OMC_OK;
}

int32_t omc_frontier_fetch_shor(omc_frontier* f, double* W, double* Xt) {
  NEED_INIT();
  if (!f || !f->big) return fail(OMC_ERR_ARG, "the frontier has no Shor rows");
  if (omcbig::big_fetch_shor(f->big, W, Xt) != 0) return fail(OMC_ERR_ARG, "%s", omcbig::big_last_error());
  return OMC_OK;
}

int32_t omc_frontier_stats(omc_frontier* f, int64_t* out8) {
  if (!f || !out8) return fail(OMC_ERR_ARG, "null argument");
  for (int q = 0; q < 8; ++q) out8[q] = 0;
  out8[0] = f->big ? OMC_ENGINE_BATCHED : OMC_ENGINE_PERSISTENT;
  if (f->big) {
    const omcbig::BigStats* s = omcbig::big_stats(f->big);
    out8[1] = s->launches; out8[2] = s->iterations; out8[3] = s->node_iterations; out8[4] = s->checks; out8[5] = s->rho_changes;
    out8[6] = (int64_t)omcbig::big_node_bytes(f->p->n, f->p->m, f->p->k, f->Lmax);
  } else {
    out8[1] = 1;
  }
  return OMC_OK;
}

int64_t omc_frontier_debug_fetch(omc_frontier* f, int32_t node, int32_t which, double* out, int64_t cap) {
  if (!f || !out || !f->big) { fail(OMC_ERR_ARG, "debug fetch needs a batched-engine frontier"); return -1; }
  return omcbig::big_debug_fetch(f->big, node, which, out, cap);
}

int32_t omc_frontier_set_tuning(omc_frontier* f, int32_t steps_max, int32_t steps_start, double track_tol, double confirm_tol) {
  if (!f) return fail(OMC_ERR_ARG, "null frontier");
  if (!f->big) return OMC_OK;
  if (steps_max > 0) f->tune.steps_max = steps_max;
  if (steps_start > 0) f->tune.steps_start = steps_start;
  if (track_tol > 0) f->tune.track_tol = track_tol;
  if (confirm_tol > 0) f->tune.confirm_tol = confirm_tol;
  return OMC_OK;
}

int32_t omc_relax_batch(omc_problem* p, int32_t B, const int32_t* node_cut_ptr, const int32_t* node_cut_ids,
                        const uint8_t* node_cut_dirs, const int32_t* warm_ids, const int32_t* save_ids,
                        const omc_relax_opts* opts, int32_t* status, double* objective, double* lower_bound,
                        int32_t* iters, double* res, double* X, double* Y, double* U, double* Theta, float* kernel_ms) {
  omc_frontier* f = nullptr;
  int32_t rc = omc_frontier_create(p, B, node_cut_ptr, node_cut_ids, node_cut_dirs, warm_ids, save_ids, &f);
  if (rc != OMC_OK) return rc;
  rc = omc_frontier_relax(f, opts, kernel_ms);
  if (rc == OMC_OK) rc = omc_frontier_fetch(f, status, objective, lower_bound, iters, res, X, Y, U, Theta);
  omc_frontier_destroy(f);
  return rc;
}

int32_t omc_objective_mse(omc_problem* p, const double* X, double* out4) {
  NEED_INIT();
  if (!p || !X || !out4) return fail(OMC_ERR_ARG, "null argument");
  const long long total = (long long)p->n * p->m;
  CU(cudaMemcpyAsync(p->Xdev.p, X, total * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  omc::objective_partial_kernel<<<g_sm_count * 4, 256, 0, g_stream>>>(p->Xdev.p, p->A.p, p->chunks.p, total, p->red.p);
  omc::objective_final_kernel<<<1, 256, 0, g_stream>>>(p->red.p, g_sm_count * 4, p->gamma, total, p->red.p + 4 * g_sm_count * 4);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out4, p->red.p + 4 * g_sm_count * 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  return OMC_OK;
}

int32_t omc_shor_score_minors(omc_problem* p, const double* Xt, int32_t n_slices, int64_t n_cand, const int32_t* cand, int64_t n_excl,
                              const int32_t* excl, int64_t n_minors, int64_t* count, int32_t* tuples, double* scores) {
  NEED_INIT();
  if (!p || !Xt || n_slices <= 0 || n_cand < 0 || n_excl < 0 || n_minors < 0 || !count || (n_cand > 0 && !cand) || (n_excl > 0 && !excl))
    return fail(OMC_ERR_ARG, "bad argument");
  if (p->n > 65535 || p->m > 65535) return fail(OMC_ERR_UNSUPPORTED, "n, m <= 65535 (16-bit minor keys)");
  *count = 0;
  if (n_cand == 0 || n_minors == 0) return OMC_OK;
  const int n = p->n, m = p->m, k = n_slices;
  for (int64_t q = 0; q < n_cand; ++q) {
    const int32_t* c = cand + 4 * q;
    if (c[0] < 0 || c[0] >= c[1] || c[1] >= n || c[2] < 0 || c[2] >= c[3] || c[3] >= m) return fail(OMC_ERR_ARG, "candidate minor %lld out of range", (long long)q);
  }
  std::vector<unsigned long long> keys((size_t)n_excl);
  for (int64_t q = 0; q < n_excl; ++q)
    keys[q] = ((unsigned long long)excl[4 * q] << 48) | ((unsigned long long)excl[4 * q + 1] << 32) | ((unsigned long long)excl[4 * q + 2] << 16) |
              (unsigned long long)excl[4 * q + 3];
  std::sort(keys.begin(), keys.end());
  DevBuf<int> dc;
  DevBuf<double> dX, ds;
  DevBuf<unsigned long long> dk;
  CU(dc.alloc((size_t)4 * n_cand)); CU(dX.alloc((size_t)k * n * m)); CU(ds.alloc((size_t)n_cand)); CU(dk.alloc(keys.empty() ? 1 : keys.size()));
  CU(cudaMemcpyAsync(dc.p, cand, (size_t)4 * n_cand * sizeof(int), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(dX.p, Xt, (size_t)k * n * m * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  if (!keys.empty()) CU(cudaMemcpyAsync(dk.p, keys.data(), keys.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, g_stream));
  omc::shor_score_kernel<<<(unsigned)((n_cand + 255) / 256), 256, 0, g_stream>>>(dc.p, n_cand, dX.p, k, n, m, dk.p, (long long)keys.size(), ds.p);
  CU(cudaGetLastError());
  std::vector<double> hs((size_t)n_cand);
  CU(cudaMemcpyAsync(hs.data(), ds.p, (size_t)n_cand * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  // top n_minors by (score, tuple), both descending: sort(...; rev = true) / partialsort!(...; rev = true) of OMC.jl:2634-2639
  std::vector<int64_t> idx;
  idx.reserve((size_t)n_cand);
  for (int64_t q = 0; q < n_cand; ++q) if (hs[q] >= 0.0) idx.push_back(q);
  auto before = [&](int64_t a_, int64_t b_) {
    if (hs[a_] != hs[b_]) return hs[a_] > hs[b_];
    for (int e = 0; e < 4; ++e) if (cand[4 * a_ + e] != cand[4 * b_ + e]) return cand[4 * a_ + e] > cand[4 * b_ + e];
    return false;
  };
  const size_t take = std::min<size_t>(idx.size(), (size_t)n_minors);
  std::partial_sort(idx.begin(), idx.begin() + take, idx.end(), before);
  *count = (int64_t)take;
  for (size_t q = 0; q < take; ++q) {
    if (tuples) for (int e = 0; e < 4; ++e) tuples[4 * q + e] = cand[4 * idx[q] + e];
    if (scores) scores[q] = hs[idx[q]];
  }
  return OMC_OK;
}

int32_t omc_profile_kernels(omc_problem* p, int32_t reps, float* out_ms) {
  NEED_INIT();
  if (!p || !out_ms || reps <= 0) return fail(OMC_ERR_ARG, "bad argument");
  const int n = p->n, m = p->m;
  const long long total = (long long)n * m;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  // [0] fused objective + MSE reduction (K9) on the staged X
  CU(cudaEventRecord(e0, g_stream));
  for (int r = 0; r < reps; ++r) {
    omc::objective_partial_kernel<<<g_sm_count * 4, 256, 0, g_stream>>>(p->Xdev.p, p->A.p, p->chunks.p, total, p->red.p);
    omc::objective_final_kernel<<<1, 256, 0, g_stream>>>(p->red.p, g_sm_count * 4, p->gamma, total, p->red.p + 4 * g_sm_count * 4);
  }
  CU(cudaEventRecord(e1, g_stream));
  CU(cudaEventSynchronize(e1));
  CU(cudaEventElapsedTime(&out_ms[0], e0, e1));
  out_ms[0] /= reps;
  // [1] mask compaction (K6): expansion, row / column counts, scans, CSR / CSC fill -- into scratch copies
  DevBuf<int> rp, cp, ci, ri;
  DevBuf<double> mk;
  CU(rp.alloc(n + 1)); CU(cp.alloc(m + 1)); CU(ci.alloc(p->nnz > 0 ? p->nnz : 1)); CU(ri.alloc(p->nnz > 0 ? p->nnz : 1)); CU(mk.alloc(total));
  const int blocks = (int)((total + 255) / 256 < g_sm_count * 8 ? (total + 255) / 256 : g_sm_count * 8);
  CU(cudaEventRecord(e0, g_stream));
  for (int r = 0; r < reps; ++r) {
    omc::mask_expand_kernel<<<blocks, 256, 0, g_stream>>>(p->chunks.p, mk.p, total);
    omc::mask_count_kernel<<<(m * 32 + 255) / 256, 256, 0, g_stream>>>(p->chunks.p, n, m, 0, cp.p);
    omc::mask_count_kernel<<<(n * 32 + 255) / 256, 256, 0, g_stream>>>(p->chunks.p, n, m, 1, rp.p);
    omc::scan_inclusive_kernel<<<1, 1024, 0, g_stream>>>(cp.p, m);
    omc::scan_inclusive_kernel<<<1, 1024, 0, g_stream>>>(rp.p, n);
    omc::mask_fill_kernel<<<(m * 32 + 255) / 256, 256, 0, g_stream>>>(p->chunks.p, n, m, 0, cp.p, ri.p);
    omc::mask_fill_kernel<<<(n * 32 + 255) / 256, 256, 0, g_stream>>>(p->chunks.p, n, m, 1, rp.p, ci.p);
  }
  CU(cudaEventRecord(e1, g_stream));
  CU(cudaEventSynchronize(e1));
  CU(cudaEventElapsedTime(&out_ms[1], e0, e1));
  out_ms[1] /= reps;
  CU(cudaGetLastError());
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return OMC_OK;
}

int32_t omc_debug_psd_project_batch(int32_t N, int32_t B, const double* Vin, double* P, double* lam, int32_t* sweeps,
                                    float* kernel_ms) {
  NEED_INIT();
  if (N <= 0 || B <= 0 || !Vin || !P || !lam || !sweeps) return fail(OMC_ERR_ARG, "bad argument");
  const omc::Geo g = omc::make_geo(N);
  if (g.NP > 104) return fail(OMC_ERR_UNSUPPORTED, "N = %d > 104", N);
  DevBuf<double> dV, dP, dL;
  DevBuf<int> dS;
  CU(dV.alloc((size_t)B * N * N)); CU(dP.alloc((size_t)B * N * N)); CU(dL.alloc((size_t)B * N)); CU(dS.alloc(B));
  CU(cudaMemcpyAsync(dV.p, Vin, (size_t)B * N * N * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  const size_t smem = ((size_t)2 * g.NP * g.ld + 2 * g.NP + 32) * 8 + (size_t)g.NP * 8 + 128;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  const int grid = B < g_sm_count ? B : g_sm_count;
  CU(cudaEventRecord(e0, g_stream));
  if (g.NP <= 32) {
    auto kern = omc::psd_project_debug_kernel<128, 8>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, g_stream>>>(N, B, dV.p, dP.p, dL.p, dS.p, 1e-13);
  } else if (g.NP <= 64) {
    auto kern = omc::psd_project_debug_kernel<256, 16>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, g_stream>>>(N, B, dV.p, dP.p, dL.p, dS.p, 1e-13);
  } else {
    auto kern = omc::psd_project_debug_kernel<512, 26>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 512, smem, g_stream>>>(N, B, dV.p, dP.p, dL.p, dS.p, 1e-13);
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(e1, g_stream));
  CU(cudaMemcpyAsync(P, dP.p, (size_t)B * N * N * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaMemcpyAsync(lam, dL.p, (size_t)B * N * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaMemcpyAsync(sweeps, dS.p, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  if (kernel_ms) CU(cudaEventElapsedTime(kernel_ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return OMC_OK;
}

int32_t omc_measure_fp64_peak(double* out2) {
  NEED_INIT();
  if (!out2) return fail(OMC_ERR_ARG, "null argument");
  DevBuf<double> d;
  const int blocks = g_sm_count * 8, threads = 256, iters = 20000;
  CU(d.alloc((size_t)blocks * threads));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  float best[2] = {1e30f, 1e30f};
  for (int rep = 0; rep < 4; ++rep) {
    float ms;
    CU(cudaEventRecord(e0, g_stream));
    omc::dfma_peak_kernel<<<blocks, threads, 0, g_stream>>>(d.p, iters);
    CU(cudaEventRecord(e1, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best[0]) best[0] = ms;
    CU(cudaEventRecord(e0, g_stream));
    omc::dmma_peak_kernel<<<blocks, threads, 0, g_stream>>>(d.p, iters / 4);
    CU(cudaEventRecord(e1, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best[1]) best[1] = ms;
  }
  CU(cudaGetLastError());
  const double fl_fma = (double)blocks * threads * iters * 8.0 * 2.0;
  const double fl_mma = (double)blocks * (threads / 32) * (iters / 4) * 8.0 * 512.0;
  out2[0] = fl_fma / (best[0] * 1e-3) * 1e-12;
  out2[1] = fl_mma / (best[1] * 1e-3) * 1e-12;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return OMC_OK;
}

}  // extern "C"

extern "C" {

int32_t omc_smallest_eigvecs_batch(int32_t n, int32_t k, int32_t B, const double* Y, const double* U, int32_t nev,
                                   double* lam, double* vec, double* breakpoint, int32_t* master_feasible) {
  NEED_INIT();
  if (n <= 0 || k <= 0 || B <= 0 || !Y || !U || !lam || !vec || !breakpoint || !master_feasible)
    return fail(OMC_ERR_ARG, "bad argument");
  if (nev != 1 && nev != 2) return fail(OMC_ERR_ARG, "nev must be 1 or 2 (OMC.jl:2467,2470)");
  if (nev > n) return fail(OMC_ERR_ARG, "nev > n");
  const omc::Geo g = omc::make_geo(n);
  DevBuf<double> dY, dU, dL, dV, dB;
  DevBuf<int> dF;
  CU(dY.alloc((size_t)B * n * n)); CU(dU.alloc((size_t)B * n * k)); CU(dL.alloc((size_t)B * nev));
  CU(dV.alloc((size_t)B * n * nev)); CU(dB.alloc((size_t)B * n)); CU(dF.alloc(B));
  CU(cudaMemcpyAsync(dY.p, Y, (size_t)B * n * n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(dU.p, U, (size_t)B * n * k * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  if (g.NP > 104) {   // beyond the in-SM eigensolver: restarted Lanczos, one CTA per node (omc_big.cuh: k_lanczos)
    const int rc = omcbig::big_smallest_eigvecs(n, k, B, dY.p, dU.p, nev, dL.p, dV.p, dB.p, dF.p, g_stream);
    if (rc != 0) return fail(rc == -4 ? OMC_ERR_UNSUPPORTED : OMC_ERR_CUDA, "%s", omcbig::big_last_error());
  } else {
    const size_t smem = omc::eigsep_smem_bytes(n);
    auto kern = omc::eigsep_kernel<256>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = B < g_sm_count * 2 ? B : g_sm_count * 2;
    kern<<<grid, 256, smem, g_stream>>>(n, k, B, dY.p, dU.p, nev, dL.p, dV.p, dB.p, dF.p);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(lam, dL.p, (size_t)B * nev * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaMemcpyAsync(vec, dV.p, (size_t)B * n * nev * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaMemcpyAsync(breakpoint, dB.p, (size_t)B * n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaMemcpyAsync(master_feasible, dF.p, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  return OMC_OK;
}

int32_t omc_altmin_batch(omc_problem* p, int32_t B, const double* U_initial, const int32_t* cut_ptr, const int32_t* cut_ids,
                         const uint8_t* cut_dirs, double eps, int32_t max_iters, double time_limit_s, double* U, double* V,
                         int32_t* converged, int32_t* n_iters, double* objectives, double* solve_time) {
  NEED_INIT();
  if (!p || !U_initial || !cut_ptr || !U || !V || !converged || !n_iters || !objectives) return fail(OMC_ERR_ARG, "null argument");
  if (B <= 0) return fail(OMC_ERR_ARG, "B must be positive");
  if (max_iters <= 0) return fail(OMC_ERR_ARG, "max_iters must be positive");
  if (p->k > omc::ALT_MAXK) return fail(OMC_ERR_UNSUPPORTED, "k = %d > %d", p->k, omc::ALT_MAXK);
  const int n = p->n, m = p->m, k = p->k;
  const int E = cut_ptr[B];
  if (cut_ptr[0] != 0 || E < 0 || (E > 0 && (!cut_ids || !cut_dirs))) return fail(OMC_ERR_ARG, "bad cut list");
  int Lmax = 0;
  for (int b = 0; b < B; ++b) {
    const int L = cut_ptr[b + 1] - cut_ptr[b];
    if (L < 0) return fail(OMC_ERR_ARG, "cut_ptr not monotone at instance %d", b);
    if (L > Lmax) Lmax = L;
  }
  for (int e = 0; e < E; ++e)
    if (cut_ids[e] < 0 || cut_ids[e] >= p->pool_size) return fail(OMC_ERR_ARG, "cut id out of range");
  const omc::AltminWs W = omc::make_altmin_ws(n, m, k, Lmax);
  const size_t ws_stride = (W.total + 1) & ~(size_t)1;
  DevBuf<double> dU, dV, dws, dobj;
  DevBuf<int> dids, dout, dptr;
  DevBuf<uint8_t> ddirs;
  CU(dU.alloc((size_t)B * n * k)); CU(dV.alloc((size_t)B * k * m)); CU(dws.alloc((size_t)B * ws_stride)); CU(dobj.alloc((size_t)B * max_iters));
  CU(dids.alloc(E > 0 ? E : 1)); CU(ddirs.alloc(E > 0 ? (size_t)E * k : 1)); CU(dout.alloc(4 * (size_t)B)); CU(dptr.alloc(B + 1));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, g_stream));
  CU(cudaMemcpyAsync(dU.p, U_initial, (size_t)B * n * k * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(dptr.p, cut_ptr, (B + 1) * sizeof(int), cudaMemcpyHostToDevice, g_stream));
  if (E > 0) {
    CU(cudaMemcpyAsync(dids.p, cut_ids, E * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    CU(cudaMemcpyAsync(ddirs.p, cut_dirs, (size_t)E * k, cudaMemcpyHostToDevice, g_stream));
  }
  CU(cudaMemsetAsync(dobj.p, 0, (size_t)B * max_iters * sizeof(double), g_stream));
  omc::AltminArgs a;
  memset(&a, 0, sizeof a);
  a.n = n; a.m = m; a.k = k; a.L = 0; a.cut_type = p->cut_type; a.fix3 = 0;
  a.gamma = p->gamma; a.eps = eps; a.inner_eps = 1e-9; a.sigma = 1e-6; a.alpha = 1.6; a.time_limit_s = time_limit_s;
  a.max_iters = max_iters; a.inner_max = 20000;
  a.A = p->A.p; a.rowptr = p->rowptr.p; a.colidx = p->colidx.p; a.colptr = p->colptr.p; a.rowidx = p->rowidx.p;
  a.pool_x = p->pool_x.p; a.pool_vhat = p->pool_vhat.p; a.cut_ids = dids.p; a.cut_dirs = ddirs.p;
  a.U = dU.p; a.V = dV.p; a.ws = dws.p; a.objectives = dobj.p; a.out_int = dout.p;
  a.cut_ptr = dptr.p; a.ws_stride = ws_stride;
  omc::altmin_kernel<1024><<<B, 1024, 0, g_stream>>>(a);
  CU(cudaGetLastError());
  int* oi = (int*)malloc(4 * (size_t)B * sizeof(int));
  if (!oi) return fail(OMC_ERR_ARG, "out of host memory");
  cudaError_t ce = cudaMemcpyAsync(U, dU.p, (size_t)B * n * k * sizeof(double), cudaMemcpyDeviceToHost, g_stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(V, dV.p, (size_t)B * k * m * sizeof(double), cudaMemcpyDeviceToHost, g_stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(objectives, dobj.p, (size_t)B * max_iters * sizeof(double), cudaMemcpyDeviceToHost, g_stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(oi, dout.p, 4 * (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, g_stream);
  if (ce == cudaSuccess) ce = cudaEventRecord(e1, g_stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(g_stream);
  float ms = 0.f;
  if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (ce != cudaSuccess) { free(oi); return fail(OMC_ERR_CUDA, "alt-min batch failed: %s", cudaGetErrorString(ce)); }
  for (int b = 0; b < B; ++b) { converged[b] = oi[4 * b]; n_iters[b] = oi[4 * b + 1]; }
  free(oi);
  if (solve_time) *solve_time = ms * 1e-3;
  return OMC_OK;
}

int32_t omc_altmin(omc_problem* p, const double* U_initial, int32_t ncuts, const int32_t* cut_ids,
                   const uint8_t* cut_dirs, double eps, int32_t max_iters, double time_limit_s, double* U, double* V,
                   int32_t* converged, int32_t* n_iters, double* objectives, double* solve_time) {
  if (ncuts < 0) return fail(OMC_ERR_ARG, "bad cut list");
  const int32_t ptr[2] = {0, ncuts};
  return omc_altmin_batch(p, 1, U_initial, ptr, cut_ids, cut_dirs, eps, max_iters, time_limit_s, U, V, converged, n_iters,
                          objectives, solve_time);
}

int32_t omc_shor_indexes(omc_problem* p, const int32_t* present_list, int32_t nlist, int64_t* count, int32_t* tuples,
                         int64_t cap, int32_t* soc, int64_t* nsoc) {
  NEED_INIT();
  if (!p || !present_list || nlist <= 0 || !count) return fail(OMC_ERR_ARG, "null argument");
  const int n = p->n, m = p->m;
  const int W = (m + 63) / 64;
  if (W > omc::SHOR_MAXW) return fail(OMC_ERR_UNSUPPORTED, "m = %d exceeds %d columns", m, 64 * omc::SHOR_MAXW);
  if (n < 2) { *count = 0; if (nsoc) *nsoc = 0; return OMC_OK; }
  int phases_h[2 * 16];
  int nphase = 0;
  if (nlist > 16) return fail(OMC_ERR_ARG, "num_entries_present list too long");
  for (int q = 0; q < nlist; ++q) {
    switch (present_list[q]) {
      case 4: phases_h[nphase++] = 0; break;
      case 3: phases_h[nphase++] = 1; break;
      case 2: phases_h[nphase++] = 2; phases_h[nphase++] = 3; break;
      case 1: phases_h[nphase++] = 4; break;
      case 0: phases_h[nphase++] = 5; break;
      default: return fail(OMC_ERR_ARG, "num_entries_present must be in 0..4");
    }
  }
  const long long npairs = (long long)n * (n - 1) / 2;
  const size_t nitems = (size_t)nphase * (size_t)npairs;
  DevBuf<unsigned long long> rowbits;
  DevBuf<long long> dcounts;
  DevBuf<int> dphases;
  CU(rowbits.alloc((size_t)n * W)); CU(dcounts.alloc(nitems)); CU(dphases.alloc(nphase));
  CU(cudaMemcpyAsync(dphases.p, phases_h, nphase * sizeof(int), cudaMemcpyHostToDevice, g_stream));
  omc::shor_rowmask_kernel<<<(n * W + 255) / 256, 256, 0, g_stream>>>(p->Mk.p, n, m, W, rowbits.p);
  omc::shor_count_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, g_stream>>>(rowbits.p, n, m, W, dphases.p, nphase, npairs, dcounts.p);
  CU(cudaGetLastError());
  long long* hc = (long long*)malloc(nitems * sizeof(long long));
  if (!hc) return fail(OMC_ERR_ARG, "out of host memory");
  cudaError_t ce = cudaMemcpyAsync(hc, dcounts.p, nitems * sizeof(long long), cudaMemcpyDeviceToHost, g_stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(g_stream);
  if (ce != cudaSuccess) { free(hc); return fail(OMC_ERR_CUDA, "Shor count failed: %s", cudaGetErrorString(ce)); }
  long long total = 0;
  for (size_t e = 0; e < nitems; ++e) { const long long c_ = hc[e]; hc[e] = total; total += c_; }   // exclusive scan, phase-major
  *count = total;
  const bool want_soc = (soc != nullptr) || (nsoc != nullptr);
  if (!tuples && !want_soc) { free(hc); return OMC_OK; }
  if (tuples && cap < total) { free(hc); return fail(OMC_ERR_ARG, "tuple buffer holds %lld, need %lld", (long long)cap, total); }
  if (total > (1ll << 28)) { free(hc); return fail(OMC_ERR_UNSUPPORTED, "%lld minors do not fit this build", total); }
  DevBuf<int> dtuples, dsoc;
  DevBuf<unsigned int> dcov;
  DevBuf<long long> dnsoc;
  const long long ncoord = (long long)n * m;
  ce = dtuples.alloc((size_t)(total > 0 ? 4 * total : 4));
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(dcounts.p, hc, nitems * sizeof(long long), cudaMemcpyHostToDevice, g_stream);
  if (ce == cudaSuccess && want_soc) {
    ce = dcov.alloc((size_t)((ncoord + 31) / 32));
    if (ce == cudaSuccess) ce = cudaMemsetAsync(dcov.p, 0, (size_t)((ncoord + 31) / 32) * sizeof(unsigned int), g_stream);
    if (ce == cudaSuccess) ce = dsoc.alloc((size_t)(2 * ncoord));
    if (ce == cudaSuccess) ce = dnsoc.alloc(1);
  }
  free(hc);
  if (ce != cudaSuccess) return fail(OMC_ERR_CUDA, "Shor allocation failed: %s", cudaGetErrorString(ce));
  if (total > 0) {
    const int wpb = 8;
    const size_t smem = (size_t)wpb * 3 * m * sizeof(int);
    long long blocks = ((long long)nitems + wpb - 1) / wpb;
    if (blocks > (long long)g_sm_count * 16) blocks = (long long)g_sm_count * 16;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(omc::shor_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    omc::shor_fill_kernel<<<(unsigned)blocks, wpb * 32, smem, g_stream>>>(rowbits.p, n, m, W, dphases.p, nphase, npairs, dcounts.p, dtuples.p,
                                                                        want_soc ? dcov.p : nullptr, m);
    CU(cudaGetLastError());
    if (tuples) CU(cudaMemcpyAsync(tuples, dtuples.p, (size_t)(4 * total) * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  }
  if (want_soc) {
    omc::shor_soc_kernel<<<1, 1024, 0, g_stream>>>(dcov.p, ncoord, n, dsoc.p, dnsoc.p);
    CU(cudaGetLastError());
    long long hn = 0;
    CU(cudaMemcpyAsync(&hn, dnsoc.p, sizeof(long long), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    if (nsoc) *nsoc = hn;
    if (soc && hn > 0) CU(cudaMemcpyAsync(soc, dsoc.p, (size_t)(2 * hn) * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  }
  CU(cudaStreamSynchronize(g_stream));
  return OMC_OK;
}

}  // extern "C"
