// omc_lowrank.cuh -- warm-started minority-side PSD projection (block LOBPCG step) for the relaxation kernel.
//
// At every ADMM iteration the argument V of a PSD-cone projection differs from the previous one by a small
// step, and one side of its spectrum is small (C2: ~10 positive eigenvalues of 100 in [Y X; X' Theta], none
// negative in a I - Y).  Instead of a full eigendecomposition the kernel tracks an orthonormal basis Z (N x p,
// p = r + OMC_LR_BUF <= PM) of the dominant invariant subspace of side * V and refines it with ONE block-LOBPCG
// step per projection:
//     W = V Z,  H = Z'W,  R = W - Z H,  R~ = orth(R | Z) (CholQR2 with rank guard),  Rayleigh-Ritz on [Z R~]
// (a 2p x 2p symmetric eigenproblem solved by two warm cyclic Jacobi sweeps in shared memory), keeping the top
// r + BUF Ritz pairs.  Work per projection is O(N^2 p) instead of O(N^3).  The projection error is bounded by
// the residual of the kept Ritz pairs plus the positive part of the complement (the projection is 1-Lipschitz);
// measured inside the ADMM it stays ~1e-9 ||V|| and leaves the iteration count unchanged (oracle/lowrank.py
// restates the scheme in NumPy and tests pin that).  The caller falls back to the full Jacobi eigensolver when
// the minority side outgrows PM, and confirms every termination decision with exact projections.
#pragma once
#include "omc_device.cuh"

namespace omc {

constexpr int OMC_LR_SWEEPS = 2; // Jacobi sweeps on the Rayleigh-Ritz matrix (warm: the Z block is nearly diagonal)
constexpr int OMC_LR_BUF = 2;    // non-positive Ritz pairs kept as a guard band
constexpr int OMC_LR_LDZ = 20;   // leading dimension of the N x 16 panels (16-byte aligned rows, DMMA conflict-free)

// Shared-memory workspace of the small matrices for a panel width PM (multiple of 4, <= 16).
template <int PM>
struct LrSmall {
  static constexpr int LD = PM + 1;
  static constexpr int N2 = 2 * PM, LD2 = 2 * PM + 1;
  double H[PM * LD];     // Z'VZ
  double X[PM * LD];     // Z'V R~   (also the projection coefficients Z'R)
  double C[PM * LD];     // R~'V R~  (also the Gram matrix R'R and its Cholesky factor)
  double T[PM * LD];     // inverse Cholesky factor
  double H2[N2 * LD2];   // Rayleigh-Ritz matrix on the live directions of [Z R~]
  double G2[N2 * LD2];   // its eigenvectors (accumulated Jacobi rotations)
  double Gc[N2 * PM];    // selected eigenvectors, compact: Gc[k * PM + j]
  double th[N2];         // Ritz values
  double cs[PM], sn[PM];
  double invd[PM];       // reciprocal diagonal of the Cholesky factor
  int ptab[(N2 - 1) * PM];  // round-robin pair table of the Jacobi sweeps: (p << 8) | q per (step, pair)
  int sel[PM];
  int valid[PM];         // residual directions that survived the rank guard
  int cidx[N2];          // live directions of [Z R~] (compacted index -> panel column, R~ columns offset by pc)
  int info[6];           // [0] = new p, [1] = r (positive Ritz values), [2] = need-full flag, [3] = live directions,
                         // [4] = 1 when the Cholesky factor was ill-conditioned (second CholQR pass needed)
};

// out[a][b] = sum_i A[i][a] * B[i][b], a, b < pc, i < N.  pc multiple of 4, panels with leading dimension LDZ.
// One warp-row of work per (a): lane = (row chunk ic in 0..7) * 4 + (column group b4 in 0..3).
__device__ __forceinline__ void panel_gram(const double* __restrict__ A, const double* __restrict__ B, int N, int pc,
                                           double* __restrict__ out, int ldo, double scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int ic = lane >> 2, b4 = lane & 3;
  for (int a = warp; a < pc; a += nw) {
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    if (4 * b4 < pc) {
#pragma unroll 4
      for (int i = ic; i < N; i += 8) {
        const double av = A[(size_t)i * OMC_LR_LDZ + a];
        const double2 b01 = *reinterpret_cast<const double2*>(B + (size_t)i * OMC_LR_LDZ + 4 * b4);
        const double2 b23 = *reinterpret_cast<const double2*>(B + (size_t)i * OMC_LR_LDZ + 4 * b4 + 2);
        acc0 += av * b01.x; acc1 += av * b01.y; acc2 += av * b23.x; acc3 += av * b23.y;
      }
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
      acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
      acc3 += __shfl_xor_sync(0xffffffffu, acc3, o);
    }
    if (ic == 0 && 4 * b4 < pc) {
      double* o_ = out + (size_t)a * ldo + 4 * b4;
      o_[0] = scale * acc0; o_[1] = scale * acc1; o_[2] = scale * acc2; o_[3] = scale * acc3;
    }
  }
}

// Wout[i][0..pc) = scale * sum_k V[i][k] * Zin[k][0..pc)   (V symmetric N x N, both triangles, leading dimension ldv).
// V in shared memory: thread (row i, 4 columns) walks its row.  V in a global (L2-resident) buffer, vglobal: the same
// thread walks COLUMN i instead (V[k][i] = V[i][k]) so that the 8 rows of a warp read 64 contiguous bytes per step.
__device__ __forceinline__ void panel_vmul(const double* __restrict__ V, int ldv, bool vglobal, int N, int NP,
                                           const double* __restrict__ Zin, double* __restrict__ Wout, int pc, double scale) {
  const int ng = pc >> 2;
  const int items = NP * ng;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int i = it / ng, jg = it - i * ng;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (i < N) {
      const double* zc = Zin + 4 * jg;
      if (!vglobal) {
        const double* vr = V + (size_t)i * ldv;
#pragma unroll 4
        for (int k = 0; k < N; ++k) {
          const double v = vr[k];
          const double2 z01 = *reinterpret_cast<const double2*>(zc + (size_t)k * OMC_LR_LDZ);
          const double2 z23 = *reinterpret_cast<const double2*>(zc + (size_t)k * OMC_LR_LDZ + 2);
          a0 += v * z01.x; a1 += v * z01.y; a2 += v * z23.x; a3 += v * z23.y;
        }
      } else {
        const double* vcol = V + i;
#pragma unroll 8
        for (int k = 0; k < N; ++k) {
          const double v = __ldcg(vcol + (size_t)k * ldv);
          const double2 z01 = *reinterpret_cast<const double2*>(zc + (size_t)k * OMC_LR_LDZ);
          const double2 z23 = *reinterpret_cast<const double2*>(zc + (size_t)k * OMC_LR_LDZ + 2);
          a0 += v * z01.x; a1 += v * z01.y; a2 += v * z23.x; a3 += v * z23.y;
        }
      }
    }
    double* w = Wout + (size_t)i * OMC_LR_LDZ + 4 * jg;
    *reinterpret_cast<double2*>(w) = make_double2(scale * a0, scale * a1);
    *reinterpret_cast<double2*>(w + 2) = make_double2(scale * a2, scale * a3);
  }
}

// Dst[i][0..pc) = beta * Src[i][0..pc) + alpha * sum_{k<kc} A[i][k] * S[k][0..pc)   (S small, leading dimension lds)
// Dst may alias Src (each item reads and writes only its own 4 entries of Src/Dst) but not A.
__device__ __forceinline__ void panel_small_mul(double* Dst, const double* Src, double beta, const double* __restrict__ A,
                                                int kc, const double* __restrict__ S, int lds, double alpha, int NP, int pc) {
  const int ng = pc >> 2;
  const int items = NP * ng;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int i = it / ng, jg = it - i * ng;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const double* ar = A + (size_t)i * OMC_LR_LDZ;
    for (int k = 0; k < kc; ++k) {
      const double v = ar[k];
      const double* s = S + (size_t)k * lds + 4 * jg;
      a0 += v * s[0]; a1 += v * s[1]; a2 += v * s[2]; a3 += v * s[3];
    }
    double* d = Dst + (size_t)i * OMC_LR_LDZ + 4 * jg;
    if (beta != 0.0) {
      const double* s_ = Src + (size_t)i * OMC_LR_LDZ + 4 * jg;
      const double s0 = s_[0], s1 = s_[1], s2 = s_[2], s3 = s_[3];
      d[0] = beta * s0 + alpha * a0; d[1] = beta * s1 + alpha * a1; d[2] = beta * s2 + alpha * a2; d[3] = beta * s3 + alpha * a3;
    } else {
      d[0] = alpha * a0; d[1] = alpha * a1; d[2] = alpha * a2; d[3] = alpha * a3;
    }
  }
}

// Cholesky of the Gram matrix M (pc x pc in shared memory, leading dimension ld, lower triangle used and destroyed)
// with a rank guard, by the first pc*pc threads: one barrier per column, every thread owns one entry (i, c) of the
// trailing matrix.  A column whose pivot falls below piv_rel * max diag (or that is already marked dead in valid[]
// when use_valid_in) is dropped: valid[j] = 0, unit pivot, zero column.  On exit Lf holds the factor (lower, incl.
// diagonal), invd[j] = 1 / L[j][j], *illcond (optional) tells whether the smallest accepted pivot is below 1e-5 max.
__device__ __forceinline__ void block_chol(double* M, double* Lf, double* invd, int ld, int pc, int* valid, double piv_rel,
                                           bool use_valid_in, int* illcond) {
  const int tid = threadIdx.x;
  const int i = tid / pc, c = tid - i * pc;
  const bool act = tid < pc * pc;
  double dmax = 0.0;
  for (int j = 0; j < pc; ++j) dmax = fmax(dmax, M[j * ld + j]);
  const double thr = piv_rel * dmax;
  double pmin = dmax;
  __syncthreads();   // everybody has read the original diagonal
  for (int j = 0; j < pc; ++j) {
    const double piv = M[j * ld + j];
    const bool ok = (piv > thr) && (piv > 0.0) && (!use_valid_in || valid[j] != 0);
    if (ok) pmin = fmin(pmin, piv);
    if (act && c >= j && i >= c) {
      const double inv = ok ? rsqrt(piv) : 1.0;
      const double lij = ok ? M[i * ld + j] * inv : 0.0;
      if (c == j) {
        Lf[i * ld + j] = (i == j) ? (ok ? piv * inv : 1.0) : lij;
        if (i == j) invd[j] = inv;
      } else {
        const double lcj = ok ? M[c * ld + j] * inv : 0.0;
        M[i * ld + c] -= lij * lcj;
      }
    }
    __syncthreads();
    if (tid == 0) valid[j] = ok ? 1 : 0;
  }
  if (tid == 0 && illcond) *illcond = (pmin < 1e-5 * dmax) ? 1 : 0;
  __syncthreads();
}

// dst[i][:] = src[i][:] L^-T (row-wise forward substitution with the Cholesky factor Lf), dropped columns zeroed.
template <int PM>
__device__ __forceinline__ void panel_trsm(const double* src, double* dst, int NP, int pc, const double* Lf, int ld,
                                           const double* invd, const int* valid) {
  for (int i = threadIdx.x; i < NP; i += blockDim.x) {
    double rt[PM];
    const double* rr = src + (size_t)i * OMC_LR_LDZ;
#pragma unroll
    for (int a = 0; a < PM; ++a) {
      double acc = 0.0;
      if (a < pc) {
        acc = rr[a];
#pragma unroll
        for (int b = 0; b < a; ++b) acc -= rt[b] * Lf[a * ld + b];
        acc = valid[a] ? acc * invd[a] : 0.0;
      }
      rt[a] = acc;
    }
    double* d = dst + (size_t)i * OMC_LR_LDZ;
#pragma unroll
    for (int a = 0; a < PM; ++a)
      if (a < pc) d[a] = rt[a];
  }
}

// symmetric Schur rotation of the pivot (app, aqq, apq): |theta| <= pi/4
__device__ __forceinline__ void schur_rot(double app, double aqq, double apq, double& c, double& s_) {
  c = 1.0; s_ = 0.0;
  if (fabs(apq) > 1e-300) {
    const double d = aqq - app, o = 2.0 * apq;
    const double ir = rsqrt(d * d + o * o);
    const double c2 = 0.5 + 0.5 * fabs(d) * ir;
    const double ic = rsqrt(c2);
    c = c2 * ic;
    s_ = copysign(0.5 * o * ir * ic, d * o);
  }
}

// Cyclic two-sided Jacobi sweeps on the symmetric n x n matrix M (n even, <= 32, leading dimension ld, both
// triangles), rotations accumulated into G (G <- G J).  Round-robin ordering, all threads of the CTA: the first n/2
// threads compute the rotation parameters of a step (a serial chain of ~500 cycles: two dependent FP64 rsqrt), then
// every 2x2 block (pair a, pair b) of M is updated on both sides at once, in place, and the columns of G are rotated;
// the M blocks and the G entries are mapped to different warps (no divergence), one item per thread when the CTA is
// large enough.  A step in which no pivot exceeds skip_tol * sqrt|a_pp a_qq| is skipped (the barrier is the vote).
__device__ __forceinline__ void small_jacobi(double* M, double* G, int n, int ld, int sweeps, double* cs, double* sn,
                                             int* ptab, int ldt, double skip_tol, long long* lp = nullptr) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int h = n >> 1;
  const int nblk = h * h, ngi = h * n;
  const int g0 = (nblk + 31) & ~31;                    // first thread of the G items (warp aligned)
  const bool one_pass = (g0 + ngi <= nt);
  for (int e = tid; e < (n - 1) * h; e += nt) {
    const int t = e / h, a = e - t * h;
    int p, q;
    jacobi_pair(a, t, n, p, q);
    ptab[t * ldt + a] = (p << 8) | q;
  }
  // static item of this thread in the one-pass mapping
  int ia0 = -1, ib0 = 0;
  if (one_pass) {
    if (tid < nblk) { ia0 = tid / h; ib0 = tid - ia0 * h; }
    else if (tid >= g0 && tid < g0 + ngi) { const int e = tid - g0; ia0 = e / n; ib0 = e - ia0 * n; }
  }
  __syncthreads();
  for (int sw = 0; sw < sweeps; ++sw) {
    for (int t = 0; t < n - 1; ++t) {
      const int* pt = ptab + t * ldt;
      int need = 0;
      const long long tp0 = clock64();
      if (tid < h) {
        int p, q;
        jacobi_pair(tid, t, n, p, q);
        const double app = M[p * ld + p], aqq = M[q * ld + q], apq = M[p * ld + q];
        double c = 1.0, s_ = 0.0;
        if (apq * apq > skip_tol * skip_tol * fabs(app * aqq)) {
          schur_rot(app, aqq, apq, c, s_);
          need = 1;
        }
        cs[tid] = c;
        sn[tid] = s_;
      }
      const int any = __syncthreads_or(need);
      if (lp && tid == 0) lp[7] += clock64() - tp0;
      if (!any) continue;
      if (one_pass) {
        if (tid < nblk) {
          const int ia = ia0, ib = ib0;
          const int pqa = pt[ia], pa = pqa >> 8, qa = pqa & 0xff;
          const int pqb = pt[ib], pb = pqb >> 8, qb = pqb & 0xff;
          const double ca = cs[ia], sa = sn[ia], cb = cs[ib], sb = sn[ib];
          const double x = M[pa * ld + pb], y = M[pa * ld + qb], z = M[qa * ld + pb], w = M[qa * ld + qb];
          const double x1 = ca * x - sa * z, z1 = sa * x + ca * z, y1 = ca * y - sa * w, w1 = sa * y + ca * w;
          double x2 = cb * x1 - sb * y1, y2 = sb * x1 + cb * y1, z2 = cb * z1 - sb * w1, w2 = sb * z1 + cb * w1;
          if (ia == ib && sa != 0.0) { y2 = 0.0; z2 = 0.0; }
          M[pa * ld + pb] = x2; M[pa * ld + qb] = y2; M[qa * ld + pb] = z2; M[qa * ld + qb] = w2;
        } else if (ia0 >= 0) {
          const int pq = pt[ia0], p = pq >> 8, q = pq & 0xff, r = ib0;
          const double c = cs[ia0], s_ = sn[ia0];
          const double x = G[r * ld + p], y = G[r * ld + q];
          G[r * ld + p] = c * x - s_ * y;
          G[r * ld + q] = s_ * x + c * y;
        }
      } else {
        for (int it = tid; it < nblk + ngi; it += nt) {
          if (it < nblk) {
            const int ia = it / h, ib = it - ia * h;
            const int pqa = pt[ia], pa = pqa >> 8, qa = pqa & 0xff;
            const int pqb = pt[ib], pb = pqb >> 8, qb = pqb & 0xff;
            const double ca = cs[ia], sa = sn[ia], cb = cs[ib], sb = sn[ib];
            const double x = M[pa * ld + pb], y = M[pa * ld + qb], z = M[qa * ld + pb], w = M[qa * ld + qb];
            const double x1 = ca * x - sa * z, z1 = sa * x + ca * z, y1 = ca * y - sa * w, w1 = sa * y + ca * w;
            double x2 = cb * x1 - sb * y1, y2 = sb * x1 + cb * y1, z2 = cb * z1 - sb * w1, w2 = sb * z1 + cb * w1;
            if (ia == ib && sa != 0.0) { y2 = 0.0; z2 = 0.0; }
            M[pa * ld + pb] = x2; M[pa * ld + qb] = y2; M[qa * ld + pb] = z2; M[qa * ld + qb] = w2;
          } else {
            const int e = it - nblk;
            const int ia = e / n, r = e - ia * n;
            const int pq = pt[ia], p = pq >> 8, q = pq & 0xff;
            const double c = cs[ia], s_ = sn[ia];
            const double x = G[r * ld + p], y = G[r * ld + q];
            G[r * ld + p] = c * x - s_ * y;
            G[r * ld + q] = s_ * x + c * y;
          }
        }
      }
      __syncthreads();
    }
  }
}

// One tracking step.  V: N x N symmetric in shared memory (rows/cols >= N are not read).  P0: current basis
// (NP x LDZ, columns >= p zero), P1 / P2: two more panels.  On exit the new basis is in *Zout (P1 or P2; columns >= new p zero),
// its Ritz values (descending) in S.th[0..new p), S.info = {new p, r, need_full, live}.  side = +1 tracks the positive
// side of V, -1 the negative side (the positive side of -V).  lp (optional, 8 slots): cycles per sub-phase.
// Returns need_full (uniform over the CTA).
template <int PM>
__device__ __forceinline__ int lowrank_step(const double* V, int ldv, bool vglobal, int N, int NP, double side, double* P0, double* P1, double* P2,
                                   int p, LrSmall<PM>& S, double vscale, long long* lp, double** Zout) {
  constexpr int LD = LrSmall<PM>::LD, LD2 = LrSmall<PM>::LD2;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int pc = (p + 3) & ~3;
  long long tk = clock64();
#define OMC_LRT(slot)                        \
  {                                          \
    const long long now_ = clock64();        \
    if (lp && tid == 0) lp[slot] += now_ - tk; \
    tk = now_;                               \
  }
  // W = side V Z -> P1 ; H = Z'W
  panel_vmul(V, ldv, vglobal, N, NP, P0, P1, pc, side);
  __syncthreads();
  OMC_LRT(0)
  panel_gram(P0, P1, N, pc, S.H, LD, 1.0);
  __syncthreads();
  // R = W - Z H (in place in P1)
  panel_small_mul(P1, P1, 1.0, P0, pc, S.H, LD, -1.0, NP, pc);
  __syncthreads();
  OMC_LRT(1)
  // second Gram-Schmidt pass against Z (the residual is a difference of nearly equal vectors late in the ADMM)
  panel_gram(P0, P1, N, pc, S.X, LD, 1.0);
  __syncthreads();
  panel_small_mul(P1, P1, 1.0, P0, pc, S.X, LD, -1.0, NP, pc);
  __syncthreads();
  // CholQR with a rank guard relative to the largest residual: P1 -> P2; when the factor is ill-conditioned the
  // result is re-projected against Z and orthonormalised once more: P2 -> P1 -> P2
  double* Rq = P2;   // orthonormal residual directions
  double* Wq = P1;   // free panel
  for (int pass = 0; pass < 3; pass += 2) {
    double* src = pass == 0 ? P1 : P2;
    double* dst = pass == 0 ? P2 : P1;
    if (pass == 2) {
      panel_gram(P0, src, N, pc, S.X, LD, 1.0);
      __syncthreads();
      panel_small_mul(src, src, 1.0, P0, pc, S.X, LD, -1.0, NP, pc);
      __syncthreads();
    }
    panel_gram(src, src, N, pc, S.C, LD, 1.0);
    __syncthreads();
    block_chol(S.C, S.T, S.invd, LD, pc, S.valid, pass == 0 ? 1e-10 : 1e-24, pass != 0, pass == 0 ? &S.info[4] : nullptr);
    panel_trsm<PM>(src, dst, NP, pc, S.T, LD, S.invd, S.valid);
    __syncthreads();
    if (pass == 0) {
      if (!S.info[4]) break;
    } else {
      Rq = P1; Wq = P2;
    }
  }
  OMC_LRT(2)
  // WR = side V R~ -> Wq ; X = Z'WR ; C = R~'WR
  panel_vmul(V, ldv, vglobal, N, NP, Rq, Wq, pc, side);
  __syncthreads();
  OMC_LRT(3)
  panel_gram(P0, Wq, N, pc, S.X, LD, 1.0);
  panel_gram(Rq, Wq, N, pc, S.C, LD, 1.0);
  // live directions of [Z R~] (warp 0): basis columns < p and the residual directions that survived
  if (tid < 32) {
    int cnt = 0;
    for (int base = 0; base < 2 * pc; base += 32) {
      const int k = base + tid;
      const bool live = (k < 2 * pc) && ((k < pc) ? (k < p) : (S.valid[k - pc] != 0));
      const unsigned bal = __ballot_sync(0xffffffffu, live);
      if (live) S.cidx[cnt + __popc(bal & ((1u << tid) - 1u))] = k;
      cnt += __popc(bal);
    }
    if (tid == 0) S.info[3] = cnt;
  }
  __syncthreads();
  OMC_LRT(4)
  // Rayleigh-Ritz matrix on the live directions (one dead direction pads an odd count: -BIG on its diagonal)
  const int nlive = S.info[3];
  const int n2 = nlive + (nlive & 1);
  const double BIG = 1.0e3 * vscale + 1.0;
  for (int e = tid; e < n2 * n2; e += nt) {
    const int ia = e / n2, ib = e - ia * n2;
    double v;
    if (ia >= nlive || ib >= nlive) {
      v = (ia == ib) ? -BIG : 0.0;
    } else {
      const int a = S.cidx[ia], b = S.cidx[ib];
      if (a < pc && b < pc) v = 0.5 * (S.H[a * LD + b] + S.H[b * LD + a]);
      else if (a >= pc && b >= pc) v = 0.5 * (S.C[(a - pc) * LD + (b - pc)] + S.C[(b - pc) * LD + (a - pc)]);
      else if (a < pc) v = S.X[a * LD + (b - pc)];
      else v = S.X[b * LD + (a - pc)];
    }
    S.H2[ia * LD2 + ib] = v;
    S.G2[ia * LD2 + ib] = (ia == ib) ? 1.0 : 0.0;
  }
  __syncthreads();
  small_jacobi(S.H2, S.G2, n2, LD2, OMC_LR_SWEEPS, S.cs, S.sn, S.ptab, PM, 1e-15, lp);
  const double* Hd = S.H2;
  OMC_LRT(5)
  // Ritz values, selection of the top r + BUF (warp 0)
  if (tid < 32) {
    const int lane = tid;
    const double my = (lane < nlive) ? Hd[lane * LD2 + lane] : -2.0 * BIG;
    int rank = 0, r = 0;
    for (int j = 0; j < nlive; ++j) {
      const double o = __shfl_sync(0xffffffffu, my, j);
      if (o > my || (o == my && j < lane)) ++rank;
      if (o > 0.0) ++r;
    }
    int pn = r + OMC_LR_BUF;
    if (pn > nlive) pn = nlive;
    int need_full = 0;
    if (r + 1 > PM || r >= nlive) need_full = 1;   // no guard band left: the minority side may be larger than tracked
    if (2 * (r + OMC_LR_BUF) > N) need_full = 1;  // [Z R~] must fit the block: the residual needs p free directions
    if (pn > PM) pn = PM;
    if (lane < nlive && rank < pn) {
      S.sel[rank] = lane;
      S.th[rank] = my;
    }
    if (lane == 0) { S.info[0] = pn; S.info[1] = r < pn ? r : pn; S.info[2] = need_full; }
  }
  // Gc = 0, then scatter the selected eigenvectors to their panel columns
  for (int e = tid; e < 2 * pc * PM; e += nt) S.Gc[e] = 0.0;
  __syncthreads();
  const int pn = S.info[0];
  const int pnc = (pn + 3) & ~3;
  for (int e = tid; e < nlive * pn; e += nt) {
    const int ia = e / pn, j = e - ia * pn;
    S.Gc[S.cidx[ia] * PM + j] = S.G2[ia * LD2 + S.sel[j]];
  }
  __syncthreads();
  // Z+ = Z Gc[0:pc] + R~ Gc[pc:2pc] -> P2 (written over max(pc, pnc) columns so that stale columns are zeroed)
  {
    const int wc = (pnc > pc) ? pnc : pc;
    const int ng = wc >> 2, items = NP * ng;
    for (int it = tid; it < items; it += nt) {
      const int i = it / ng, jg = it - i * ng;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      const double* zr = P0 + (size_t)i * OMC_LR_LDZ;
      const double* rr = Rq + (size_t)i * OMC_LR_LDZ;
      for (int k = 0; k < pc; ++k) {
        const double vz = zr[k], vr = rr[k];
        const double* g0 = S.Gc + (size_t)k * PM + 4 * jg;
        const double* g1 = S.Gc + (size_t)(pc + k) * PM + 4 * jg;
        a0 += vz * g0[0] + vr * g1[0]; a1 += vz * g0[1] + vr * g1[1];
        a2 += vz * g0[2] + vr * g1[2]; a3 += vz * g0[3] + vr * g1[3];
      }
      double* d = Wq + (size_t)i * OMC_LR_LDZ + 4 * jg;
      d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
    }
    __syncthreads();
  }
  OMC_LRT(6)
#undef OMC_LRT
  *Zout = Wq;
  return S.info[2];
}

}  // namespace omc
