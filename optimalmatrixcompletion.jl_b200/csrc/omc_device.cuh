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
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }
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

// In-place Gauss-Jordan inversion of the SPD r x r matrix M (leading dimension r), no pivoting.
__device__ inline void spd_invert(double* M, int r) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int p = 0; p < r; ++p) {
    __syncthreads();
    const double inv = 1.0 / M[(size_t)p * r + p];
    __syncthreads();
    for (int j = tid; j < r; j += nt)
      if (j != p) M[(size_t)p * r + j] *= inv;
    __syncthreads();
    for (int e = tid; e < r * r; e += nt) {
      const int i = e / r, j = e - i * r;
      if (i != p && j != p) M[e] -= M[(size_t)i * r + p] * M[(size_t)p * r + j];
    }
    __syncthreads();
    for (int i = tid; i < r; i += nt)
      if (i != p) M[(size_t)i * r + p] *= -inv;
    if (tid == 0) M[(size_t)p * r + p] = inv;
  }
  __syncthreads();
}


// ------------------------------------------------------------------------------------------------
// Geometry of one symmetric block held in shared memory: N real rows, NP = N rounded up to 8 (DMMA
// tile), leading dimension ld = NP + 1.  An ODD leading dimension makes a warp's 64-bit accesses
// conflict-free both along a row (consecutive doubles) and down a column (stride ld: 16 lanes hit 16
// distinct 8-byte banks), which the Jacobi column rotations need; the price is a 2-way conflict on part
// of the DMMA fragment loads (4 rows x 4 doubles), ~10 % of the kernel.
// ------------------------------------------------------------------------------------------------
struct Geo {
  int N, NP, ld;
};
__host__ __device__ inline Geo make_geo(int N) {
  Geo g;
  g.N = N;
  g.NP = (N + 7) & ~7;
  g.ld = g.NP + 1;
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
__device__ __noinline__ void gemm_rows_inplace(double* M, const double* Q, int NP, int ld) {
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
__device__ __noinline__ void gemm_cols_inplace(double* M, const double* Q, int NP, int ld) {
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
//
// After every sweep the remaining off-diagonal mass is measured directly (an N^2 pass, ~1% of a sweep);
// the solver stops when off(S) <= tol * ||S||_F or after max_sweeps.
//
// projection_mode: the caller only needs the projection onto the PSD cone, i.e. the split into the
// positive and the negative invariant subspace and the eigenpairs of the SMALLER side.  Indices whose
// diagonal entry dominates its row by a factor 2 (|s_ii| > 2 sum_j |s_ij|) are "safe": by Gershgorin the
// principal submatrix on the safe indices of one sign is definite with that sign, whatever basis spans it.
// Rotations between two safe indices of the majority sign are therefore skipped (inside degenerate
// clusters of the dual they would be O(1) rotations at every ADMM iteration, for nothing); every other pair
// is rotated to tolerance, so each non-skipped index ends up as an exact eigenpair decoupled from the rest.
// *skip_sign returns 0 (nothing skipped), +1 (safe positive indices skipped: use the negative side) or -1;
// skip[i] != 0 marks the skipped indices.  With projection_mode = 0 this is a full eigensolver.
//
// cs / sn: NP/2 doubles; rot: 3*NP/2 ints (>= 1 + NP/2 used); skip: NP ints; scratch: 32 doubles.
// Returns the sweep count.
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

__device__ __noinline__ int jacobi_sym(double* S, double* Q, int NP, int ld, double tol, int max_sweeps, double* cs,
                                 double* sn, int* rot, double* scratch, int projection_mode = 0, int* skip = nullptr,
                                 int* skip_sign = nullptr, double* stats = nullptr) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int H = NP >> 1;
  // 2-D thread map without integer division in the hot loops: tx = pair j (fastest), ty = row group
  const int TX = (H <= 16) ? 16 : ((H <= 32) ? 32 : 64);
  const int tx = tid & (TX - 1), ty = tid / TX, TY = nt / TX;
  const bool pm = projection_mode && skip != nullptr;
  double acc = 0.0;
  for (int e = tid; e < NP * NP; e += nt) {
    const int r = e / NP, c = e - r * NP;
    const double v = S[(size_t)r * ld + c];
    acc += v * v;
  }
  const double fro2 = block_sum(acc, scratch);
  if (pm && tid < NP) skip[tid] = 0;
  if (skip_sign && tid == 0) *skip_sign = 0;
  __syncthreads();
  if (fro2 == 0.0) return 0;
  const double stop2 = tol * tol * fro2;
  const double thr = 1e-20 * sqrt(fro2);
  int sweeps = 0;
  double nrot = 0.0, nskip = 0.0;
  for (;;) {
    // ---- classification (projection mode) and the off-diagonal mass that still matters
    if (pm) {
      int c = 0;
      if (tid < NP) {
        const double* row = S + (size_t)tid * ld;
        double rs = 0.0;
        for (int j = 0; j < NP; ++j) rs += fabs(row[j]);
        const double d = row[tid];
        rs -= fabs(d);
        c = (d > 2.0 * rs) ? 1 : ((d < -2.0 * rs) ? -1 : 0);
      }
      const int cpos = __syncthreads_count(c > 0);
      const int cneg = __syncthreads_count(c < 0);
      const int sg = (cpos == 0 && cneg == 0) ? 0 : ((cneg >= cpos) ? -1 : 1);
      if (tid < NP) skip[tid] = (sg != 0 && c == sg) ? 1 : 0;
      if (skip_sign && tid == 0) *skip_sign = sg;
      if (tid == 0) nskip += (sg < 0) ? cneg : ((sg > 0) ? cpos : 0);
      __syncthreads();
    }
    double off2 = 0.0;
    for (int r = ty; r < NP; r += TY) {
      const double* row = S + (size_t)r * ld;
      const int sr = pm ? skip[r] : 0;
      for (int c = tx; c < NP; c += TX)
        if (c != r && !(sr && skip[c])) off2 += row[c] * row[c];
    }
    off2 = block_sum(off2, scratch);
    if (off2 <= stop2 || sweeps >= max_sweeps) break;
    // ---- one sweep
    for (int t = 0; t < NP - 1; ++t) {
      if (tid == 0) rot[0] = 0;  // rot[0] = number of rotated pairs of this step, rot[1 + a] = (p << 16) | q
      __syncthreads();
      if (tid < H) {
        int p, q;
        jacobi_pair(tid, t, NP, p, q);
        const double apq = S[(size_t)p * ld + q];
        const bool skipped = pm && skip[p] && skip[q];
        if (!skipped && fabs(apq) > thr) {
          // symmetric Schur rotation with |theta| <= pi/4: cos 2theta = |d| / r, sin 2theta = +-o / r
          const double d = S[(size_t)q * ld + q] - S[(size_t)p * ld + p], o = 2.0 * apq;
          const double ir = rsqrt(d * d + o * o);
          const double c2 = 0.5 + 0.5 * fabs(d) * ir;       // cos^2 theta
          const double ic = rsqrt(c2);
          const double c = c2 * ic;
          const double s_ = copysign(0.5 * o * ir * ic, d * o);  // sin theta, sign of tau * ... = sign(d) sign(o)
          const int a = atomicAdd(&rot[0], 1);
          rot[1 + a] = (p << 16) | q;
          cs[a] = c;
          sn[a] = s_;
          nrot += 1.0;
        }
      }
      __syncthreads();
      const int nr = rot[0];
      if (nr > 0) {
        // pass 1: S <- J' S  (rows p, q of every rotated pair; lanes walk along the row)
        for (int e = tid; e < nr * NP; e += nt) {
          const int a = e / NP, c = e - a * NP;
          const int pq = rot[1 + a], p = pq >> 16, q = pq & 0xffff;
          const double ca = cs[a], sa = sn[a];
          double* rp_ = S + (size_t)p * ld + c;
          double* rq_ = S + (size_t)q * ld + c;
          const double x = *rp_, y = *rq_;
          *rp_ = ca * x - sa * y;
          *rq_ = sa * x + ca * y;
        }
        __syncthreads();
        // pass 2: S <- S J and Q <- Q J (columns p, q; lanes walk down the column, conflict-free for odd ld)
        for (int e = tid; e < 2 * nr * NP; e += nt) {
          const int h = e / (nr * NP), e2 = e - h * nr * NP;
          const int a = e2 / NP, r = e2 - a * NP;
          const int pq = rot[1 + a], p = pq >> 16, q = pq & 0xffff;
          const double ca = cs[a], sa = sn[a];
          double* base = (h == 0 ? S : Q) + (size_t)r * ld;
          const double x = base[p], y = base[q];
          double xn = ca * x - sa * y, yn = sa * x + ca * y;
          if (h == 0) {  // the rotated entry is annihilated exactly
            if (r == p) yn = 0.0;
            if (r == q) xn = 0.0;
          }
          base[p] = xn;
          base[q] = yn;
        }
      }
      __syncthreads();
    }
    ++sweeps;
  }
  if (stats) {  // diagnostics: rotations applied, skipped-set sizes (summed over classification passes)
    nrot = block_sum(nrot, scratch);
    if (tid == 0) {
      stats[0] += nrot;
      stats[1] += nskip;
      stats[2] += 1.0;
    }
  }
  return sweeps;
}

}  // namespace omc
