// omc_device.cuh -- device-side building blocks shared by the omc_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef OMC_USE_TMA
#define OMC_USE_TMA 1
#endif

namespace omc {

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core MMA (DMMA 8x8x4): D(8x8) = A(8x4,row) * B(4x8,col) + C.
//   lane l: a = A[l>>2][l&3], b = B[l&3][l>>2], c0/c1 = C[l>>2][2*(l&3) + {0,1}]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// ------------------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, 1-D, no tensor map) + mbarrier.  One thread issues; everybody
// waits on the mbarrier.  Sizes are multiples of 16 bytes, addresses 16-byte aligned.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "OMC_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra OMC_DONE_%=;\n"
      "bra OMC_WAIT_%=;\n"
      "OMC_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

// ------------------------------------------------------------------------------------------------
// block-wide reductions (NT threads, scratch >= 32 doubles).  Result valid in every thread.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = (lane < nw) ? scratch[lane] : 0.0;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = (lane < nw) ? scratch[lane] : -1.0e300;
  r = warp_max(r);
  return r;
}

// ------------------------------------------------------------------------------------------------
// Geometry of one symmetric block held in shared memory: N real rows, NP = N rounded up to 8 (DMMA
// tile), leading dimension ld = NP + 4 (ld % 8 == 4 makes both DMMA fragment patterns, 4 rows x 4
// consecutive doubles, hit 16 distinct 8-byte banks).
// ------------------------------------------------------------------------------------------------
struct Geo {
  int N, NP, ld;
};
__host__ __device__ inline Geo make_geo(int N) {
  Geo g;
  g.N = N;
  g.NP = (N + 7) & ~7;
  g.ld = g.NP + 4;
  return g;
}

// ------------------------------------------------------------------------------------------------
// In-place GEMMs on the shared-memory pair (M, Q), all warps of the CTA cooperating:
//   rows:  M <- M * Q      (row panel of 8 rows lives in registers as DMMA A fragments)
//   cols:  M <- Q' * M     (column panel of 8 columns lives in registers as DMMA B fragments)
// Output panel p depends only on input panel p, so the update is in place once every warp that
// shares a panel has loaded its fragments (one __syncthreads).  KMAX >= NP/4.
// ------------------------------------------------------------------------------------------------
template <int KMAX>
__device__ __forceinline__ void gemm_rows_inplace(double* M, const double* Q, int NP, int ld) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int T = NP >> 3, KS = NP >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int nsplit = (nw >= T) ? (nw / T) : 1;
  const int rounds = (nw >= T) ? 1 : (T + nw - 1) / nw;
  for (int rd = 0; rd < rounds; ++rd) {
    int p, q;
    if (nw >= T) { p = warp % T; q = warp / T; } else { p = warp + rd * nw; q = 0; }
    const bool active = (p < T) && (q < nsplit);
    double a[KMAX];
    if (active) {
      const double* row = M + (size_t)(p * 8 + g) * ld + t;
#pragma unroll
      for (int kk = 0; kk < KMAX; ++kk) a[kk] = (kk < KS) ? row[kk * 4] : 0.0;
    }
    __syncthreads();
    if (active) {
      for (int ct = q; ct < T; ct += nsplit) {
        double c0 = 0.0, c1 = 0.0;
        const double* bcol = Q + (size_t)t * ld + ct * 8 + g;
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk) {
          if (kk < KS) {
            const double b = bcol[(size_t)kk * 4 * ld];
            dmma884(c0, c1, a[kk], b, c0, c1);
          }
        }
        double* out = M + (size_t)(p * 8 + g) * ld + ct * 8 + 2 * t;
        out[0] = c0;
        out[1] = c1;
      }
    }
    __syncthreads();
  }
}

template <int KMAX>
__device__ __forceinline__ void gemm_cols_inplace(double* M, const double* Q, int NP, int ld) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int T = NP >> 3, KS = NP >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int nsplit = (nw >= T) ? (nw / T) : 1;
  const int rounds = (nw >= T) ? 1 : (T + nw - 1) / nw;
  for (int rd = 0; rd < rounds; ++rd) {
    int p, q;
    if (nw >= T) { p = warp % T; q = warp / T; } else { p = warp + rd * nw; q = 0; }
    const bool active = (p < T) && (q < nsplit);
    double b[KMAX];
    if (active) {
      const double* col = M + (size_t)t * ld + p * 8 + g;
#pragma unroll
      for (int kk = 0; kk < KMAX; ++kk) b[kk] = (kk < KS) ? col[(size_t)kk * 4 * ld] : 0.0;
    }
    __syncthreads();
    if (active) {
      for (int rt = q; rt < T; rt += nsplit) {
        double c0 = 0.0, c1 = 0.0;
        // A = Q' : A[rt*8+g][kk*4+t] = Q[kk*4+t][rt*8+g]
        const double* acol = Q + (size_t)t * ld + rt * 8 + g;
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk) {
          if (kk < KS) {
            const double a = acol[(size_t)kk * 4 * ld];
            dmma884(c0, c1, a, b[kk], c0, c1);
          }
        }
        double* out = M + (size_t)(rt * 8 + g) * ld + p * 8 + 2 * t;
        out[0] = c0;
        out[1] = c1;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Parallel cyclic two-sided Jacobi on the symmetric matrix S (NP x NP, NP even, both triangles kept)
// in shared memory, accumulating the rotations into Q (Q <- Q * J).  Round-robin ordering: step t
// pairs (NP-1, t) and ((t+i) mod (NP-1), (t-i) mod (NP-1)), i = 1 .. NP/2-1, so each sweep visits all
// NP(NP-1)/2 pairs in NP-1 steps of NP/2 disjoint rotations.
// After every sweep the remaining off-diagonal mass is measured directly (an N^2 pass, ~1% of a sweep);
// the solver stops when off(S) <= tol * ||S||_F or after max_sweeps.
// cs / sn / rot: shared arrays of NP/2 entries.  Returns the number of sweeps (same in all threads).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void jacobi_pair(int i, int t, int NP, int& p, int& q) {
  const int M = NP - 1;
  if (i == 0) {
    p = NP - 1;
    q = t;
  } else {
    p = t + i;
    if (p >= M) p -= M;
    q = t - i;
    if (q < 0) q += M;
  }
}

__device__ inline int jacobi_sym(double* S, double* Q, int NP, int ld, double tol, int max_sweeps, double* cs,
                                 double* sn, int* rot, double* scratch) {
  // rot[] holds, per pair i of the current step: rot[i] = rotate flag, rot[H + i] = p_i, rot[2H + i] = q_i
  const int tid = threadIdx.x, nt = blockDim.x;
  const int H = NP >> 1;
  int* pidx = rot + H;
  int* qidx = rot + 2 * H;
  // 2-D thread map without integer division in the hot loops: tx = pair j (fastest), ty = row group
  const int TX = (H <= 16) ? 16 : ((H <= 32) ? 32 : 64);
  const int tx = tid & (TX - 1), ty = tid / TX, TY = nt / TX;
  double acc = 0.0, acco = 0.0;
  for (int e = tid; e < NP * NP; e += nt) {
    const int r = e / NP, c = e - r * NP;
    const double v = S[(size_t)r * ld + c];
    acc += v * v;
    if (r != c) acco += v * v;
  }
  const double fro2 = block_sum(acc, scratch);
  const double offin2 = block_sum(acco, scratch);
  if (fro2 == 0.0) return 0;
  const double stop2 = tol * tol * fro2;
  if (offin2 <= stop2) return 0;  // already diagonal to the requested accuracy
  const double thr = 1e-20 * sqrt(fro2);
  int sweeps = 0;
  for (; sweeps < max_sweeps;) {
    for (int t = 0; t < NP - 1; ++t) {
      if (tid < H) {
        int p, q;
        jacobi_pair(tid, t, NP, p, q);
        const double apq = S[(size_t)p * ld + q];
        double c = 1.0, s = 0.0;
        int r = 0;
        if (fabs(apq) > thr) {
          const double app = S[(size_t)p * ld + p], aqq = S[(size_t)q * ld + q];
          const double tau = (aqq - app) / (2.0 * apq);
          const double tt = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = rsqrt(1.0 + tt * tt);
          s = tt * c;
          r = 1;
        }
        cs[tid] = c;
        sn[tid] = s;
        rot[tid] = r;
        pidx[tid] = p;
        qidx[tid] = q;
      }
      __syncthreads();
      // S <- J' S J : every 2x2 block (i, j) is owned by one thread and written in its own orientation
      // only (row-wise accesses, lanes walk consecutive columns -> no bank conflicts, no mirrored stores)
      if (tx < H) {
        const int rj = rot[tx], pj = pidx[tx], qj = qidx[tx];
        const double cj = cs[tx], sj = sn[tx];
        for (int i = ty; i < H; i += TY) {
          const int ri = rot[i];
          if (!(ri | rj)) continue;
          const int pi = pidx[i], qi = qidx[i];
          const double ci = cs[i], si = sn[i];
          double* rp_ = S + (size_t)pi * ld;
          double* rq_ = S + (size_t)qi * ld;
          const double b00 = rp_[pj], b01 = rp_[qj], b10 = rq_[pj], b11 = rq_[qj];
          const double t00 = ci * b00 - si * b10, t01 = ci * b01 - si * b11;
          const double t10 = si * b00 + ci * b10, t11 = si * b01 + ci * b11;
          double n00 = cj * t00 - sj * t01, n01 = sj * t00 + cj * t01;
          double n10 = cj * t10 - sj * t11, n11 = sj * t10 + cj * t11;
          if (i == tx) {
            n01 = 0.0;
            n10 = 0.0;
          }
          rp_[pj] = n00;
          rp_[qj] = n01;
          rq_[pj] = n10;
          rq_[qj] = n11;
        }
        // Q <- Q J
        if (rj) {
          for (int r = ty; r < NP; r += TY) {
            double* row = Q + (size_t)r * ld;
            const double xp = row[pj], xq = row[qj];
            row[pj] = cj * xp - sj * xq;
            row[qj] = sj * xp + cj * xq;
          }
        }
      }
      __syncthreads();
    }
    ++sweeps;
    double off2 = 0.0;
    for (int r = ty; r < NP; r += TY) {
      const double* row = S + (size_t)r * ld;
      for (int c = tx; c < NP; c += TX)
        if (c != r) off2 += row[c] * row[c];
    }
    off2 = block_sum(off2, scratch);
    if (off2 <= stop2) break;
  }
  return sweeps;
}

}  // namespace omc
