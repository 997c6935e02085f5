// omc_big.cuh -- the batched large-block relaxation engine (round 2): the node relaxation of
// matrix_completion_SDP_relaxation (/root/reference/src/OptimalMatrixCompletion.jl:1431-1943, "OMC.jl") for PSD blocks
// that do not fit one SM's shared memory (config 4: n + m = 200, config 5: n + m = 2000), run in LOCKSTEP over a whole
// frontier of open nodes: every ADMM iteration is a short sequence of kernels whose grids span (tile, node), so the
// machine is filled by the frontier, the state of a node streams from HBM once per pass, and nothing about a node has
// to fit on chip.  CPU restatement with the same steps and constants: oracle/bigblock.py.
//
//   v-form ADMM (COSMO/OSQP splitting of oracle/relaxation.py):  v <- v + alpha (z~ - P_K(v)),  mu = rho (v - P_K(v)).
//   The projection of a PSD-block argument V_b is never formed: the minority spectral side of V_b (positive side of
//   [Y X; X' Theta] and [Y U; U' I], negative side of aI - Y) is tracked as Z_b diag(theta_b) Z_b' with a panel of
//   PM = 32 columns (16 for aI - Y) refined by block-LOBPCG steps on [Z, R] (R = residual of V Z, explicitly re-projected against Z),
//   Rayleigh-Ritz on the 64 x 64 projected matrix by a CTA-parallel cyclic Jacobi; every pass over V_b uses V_b and the factor.
//
// Kernels (grid.y = active node, grid.z = PSD block where it applies):
//   k_xt      X and Theta regions: w-update, v-update and w relaxation fused in one pass (these entries meet no dense row)
//   k_y1      Y and U regions, pass A: diagonal part of the w-update, per-tile partial sums of the dense rows
//             (trace row, cut rows x'U_j, x'Yx)
//   k_small   per node: dense-row right-hand side, Woodbury solve (explicit inverse, rebuilt when rho changes), scalar rows
//   k_y2      Y and U regions, pass B: rank-(1 + L) correction, v-update of the three blocks' Y / U parts, w relaxation
//   k_prod    W = side * V_b P for a 16-column panel P (FP64 DMMA m8n8k4, V streamed once), partial Gram matrices in the epilogue
//   k_resid1/2, k_rr, k_update   the rest of a tracker step
//   k_check + k_decide   residuals, dual objective, certified bound, termination / cut-off / infeasibility, rho adaptation
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/omc_b200.h"

namespace omcbig {

// FP64 tensor-core MMA (DMMA 8x8x4): D(8x8) = A(8x4,row) * B(4x8,col) + C.
//   lane l: a = A[l>>2][l&3], b = B[l&3][l>>2], c0/c1 = C[l>>2][2*(l&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
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
// block-wide reductions (scratch >= 32 doubles); result valid in every thread
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = (lane < nw) ? scratch[lane] : 0.0;
  return warp_sum(r);
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = (lane < nw) ? scratch[lane] : -1.0e300;
  return warp_max(r);
}

constexpr int PM = 32;        // tracker panel width (blocks 1 and 2)
constexpr int P3 = 16;        // panel width of block 3 (negative side of aI - Y: a handful of eigenvalues)
constexpr int TS = 64;        // region tile
constexpr int NCHK = 16;      // per-tile partials of the residual check
constexpr int ISTR = 32;      // ints per node
constexpr int SSTR = 48;      // scalar doubles per node
// int slots
enum { I_STATUS = 0, I_L = 1, I_CONFIRM = 2, I_WASCONF = 3, I_MORE = 4 /*3*/, I_Q = 7 /*3*/, I_R = 10 /*3*/, I_ITERS = 13, I_DONE = 14,
       I_ADAPTED = 15, I_SLOT = 16 };
// scalar slots
enum { S_RHO = 0, S_V4 = 1, S_RP = 2, S_RD = 3, S_OBJP = 4, S_OBJD = 5, S_LB = 6, S_CFAC = 7, S_RES = 8 /*3*/, S_NP = 11, S_ND = 12,
       S_RESID2 = 13 /*3*/, S_HS2 = 16 /*3*/ };

struct Layout {
  int n, m, k, N[3], p[3], Lcap, rcap, tn, tm, nt[3];   // nt[b] = row tiles of block b
  int tilesXT, tilesYU, tilesAll;
  size_t X, Y, T, U, V[3], Yt, Ut, v5, vv, vg, xs, clb, cub, cal, cbe, G, Minv, rhs, cw;
  size_t Z[3], W[3], R[3], W2[3], th[3], H[3], Q[3], partA[3], partB[3], rows_part, chk_part, scal, total;
};

inline size_t al4(size_t o) { return (o + 3) & ~(size_t)3; }

inline Layout make_layout(int n, int m, int k, int Lcap) {
  Layout L;
  L.n = n; L.m = m; L.k = k; L.Lcap = Lcap; L.rcap = 1 + Lcap * (k + 1);
  L.N[0] = n + m; L.N[1] = n + k; L.N[2] = n;
  for (int b = 0; b < 3; ++b) {
    const int cap = (b == 2) ? P3 : PM;
    L.p[b] = L.N[b] < cap ? L.N[b] : cap;   // N <= cap: the panel is a complete eigenbasis and the projection is exact
    L.nt[b] = (L.N[b] + TS - 1) / TS;
  }
  L.tn = (n + TS - 1) / TS; L.tm = (m + TS - 1) / TS;
  L.tilesXT = L.tn * L.tm + L.tm * (L.tm + 1) / 2;
  L.tilesYU = L.tn * (L.tn + 1) / 2 + L.tn;
  L.tilesAll = L.tilesXT + L.tilesYU;
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t r = o; o = al4(o + cnt); return r; };
  L.X = take((size_t)n * m); L.Y = take((size_t)n * n); L.T = take((size_t)m * m); L.U = take((size_t)n * k);
  for (int b = 0; b < 3; ++b) L.V[b] = take((size_t)L.N[b] * L.N[b]);
  L.Yt = take((size_t)n * n); L.Ut = take((size_t)n * k);
  L.v5 = take((size_t)n * k); L.vv = take((size_t)Lcap * k + 1); L.vg = take((size_t)Lcap + 1);
  L.xs = take((size_t)Lcap * n + 1); L.clb = take((size_t)Lcap * k + 1); L.cub = take((size_t)Lcap * k + 1);
  L.cal = take((size_t)Lcap * k + 1); L.cbe = take((size_t)Lcap + 1);
  L.G = take((size_t)L.rcap * L.rcap); L.Minv = take((size_t)L.rcap * L.rcap); L.rhs = take(L.rcap); L.cw = take(L.rcap);
  for (int b = 0; b < 3; ++b) {
    L.Z[b] = take((size_t)L.N[b] * PM); L.W[b] = take((size_t)L.N[b] * PM); L.R[b] = take((size_t)L.N[b] * PM);
    L.W2[b] = take((size_t)L.N[b] * PM); L.th[b] = take(PM); L.H[b] = take(PM * PM); L.Q[b] = take(2 * PM * PM);
    L.partA[b] = take((size_t)L.nt[b] * PM * PM); L.partB[b] = take((size_t)L.nt[b] * PM * PM);
  }
  L.rows_part = take((size_t)L.tilesYU * L.rcap);
  L.chk_part = take((size_t)L.tilesAll * NCHK);
  L.scal = take(SSTR);
  L.total = o;
  return L;
}

constexpr int B5 = 15;
// ---- per-node Shor record (doubles), stride SL.total ---------------------------------------------------------------
struct ShorLayout {
  int k, npair, K9;          // K9 = (k+1)(k+2)/2 packed entries of the (k+1) x (k+1) block
  long long C, nm, nv1, nv2;
  size_t Xt, Wd, H, V1, V2, V3, Xtt, Wdt, Ht, V1t, V2t, V3t, GX, GW, GH, vB, TB, v9, T9, v6, v7, vs, Ts, gTd, Fd, chk, total;
};
inline ShorLayout make_shor_layout(int n, int m, int k, long long nm, long long nv1, long long nv2) {
  ShorLayout S;
  S.k = k; S.npair = k * (k - 1) / 2; S.K9 = (k + 1) * (k + 2) / 2;
  S.C = (long long)n * m; S.nm = nm; S.nv1 = nv1; S.nv2 = nv2;
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t r = o; o = al4(o + cnt); return r; };
  const size_t C = (size_t)S.C;
  S.Xt = take(k * C); S.Wd = take(k * C); S.H = take((size_t)S.npair * C + 1);
  S.V1 = take((size_t)k * nv1 + 1); S.V2 = take((size_t)k * nv2 + 1); S.V3 = take((size_t)k * nm + 1);
  S.Xtt = take(k * C); S.Wdt = take(k * C); S.Ht = take((size_t)S.npair * C + 1);
  S.V1t = take((size_t)k * nv1 + 1); S.V2t = take((size_t)k * nv2 + 1); S.V3t = take((size_t)k * nm + 1);
  S.GX = take(k * C); S.GW = take(k * C); S.GH = take((size_t)S.npair * C + 1);
  S.vB = take((size_t)k * nm * B5 + 1); S.TB = take((size_t)k * nm * B5 + 1);
  S.v9 = take(C * S.K9); S.T9 = take(C * S.K9);
  S.v6 = take(m); S.v7 = take(k * C); S.vs = take(C * 3); S.Ts = take(C * 3);
  S.gTd = take(m); S.Fd = take(m); S.chk = take(64 * 8);
  S.total = o;
  return S;
}

struct ShorDev {          // problem-level structure (shared by all nodes) + per-node records
  int on;
  ShorLayout SL;
  double* SS;             // node Shor records
  const int* minors;      // [nm][4]
  const int* mv;          // [nm][4]: V1 id (row i1), V1 id (row i2), V2 id (col j1), V2 id (col j2)
  const int* cptr; const int* cinc;     // coordinate incidence: entries minor * 4 + slot
  const int* v1ptr; const int* v1inc;   // V1 incidence: minor * 2 + which (0: entry (2,1), 1: entry (4,3))
  const int* v2ptr; const int* v2inc;   // V2 incidence: minor * 2 + which (0: entry (3,1), 1: entry (4,2))
  const unsigned char* flags;           // per coordinate: bit 0 covered, bit 1 SOC
  const int* cnt;                       // minors per coordinate
};


struct Opts {
  double eps_abs, eps_rel, sigma, alpha, rho0, cutoff, track_tol, confirm_tol, adapt_thresh;
  int max_iter, check_every, adapt_every, steps_max, steps_start, fix_linear3_right, cut_type, infeasible_by_bound, jacobi_sweeps, window;
};

struct BigArgs {
  Layout L;
  double* S;            // node records, stride L.total
  int* I;               // node ints, stride ISTR
  const int* active;    // active node slots
  const double* AM;     // row-major n x m: mask * A
  const unsigned char* Mk;  // row-major n x m: mask
  double a, sa, cT, c0, ktr;
  Opts o;
  int it;               // current iteration (1-based)
  int step;             // tracker step within the iteration
  int force;            // this iteration is followed by an off-schedule check (confirm pending somewhere)
  ShorDev sh;           // Shor valid-inequality rows (omc_big_shor.cuh); sh.on = 0 without them
};

__device__ __forceinline__ double* node_ptr(const BigArgs& a, int slot) { return a.S + (size_t)slot * a.L.total; }
__device__ __forceinline__ int* node_int(const BigArgs& a, int slot) { return a.I + (size_t)slot * ISTR; }

// ------------------------------------------------------------------------------------------------------------------
// tile helpers: 64 x 64 tile, 256 threads; thread (ty, tx) owns entries (ty + 16 r, tx + 16 c), r, c = 0..3, so that a
// half-warp reads / writes 16 consecutive doubles of a row.
// ------------------------------------------------------------------------------------------------------------------
constexpr int ZLD = PM + 1;

__device__ __forceinline__ double hash_unit(unsigned long long i, unsigned long long j, unsigned long long seed) {
  unsigned long long h = i * 0x9E3779B97F4A7C15ull + j * 0xC2B2AE3D27D4EB4Full + seed * 0x165667B19E3779F9ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// L2 prefetch of the 64 x 64 tile (i0, j0) of a row-major array in the thread-to-entry mapping of the region kernels (one lane per
// 32-byte sector).  Issued before the panels are staged and the low-rank tile is computed, so that the tile's DRAM latency
// overlaps that FP64 work instead of stalling the element loop: a prefetch holds no register.
__device__ __forceinline__ void prefetch_tile(const double* __restrict__ base, size_t ld, int i0, int j0, int rows, int cols, int ty, int tx) {
  if (tx & 3) return;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + ty + 16 * r, j = j0 + tx + 16 * c;
      if (i < rows && j < cols) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)i * ld + j));
    }
}

// stage the first nc columns of rows [row0, row0 + 64) of a panel (global ld PM) into shared memory (ld), column a scaled
// by sc[a] (sqrt(theta+) -> the low-rank term becomes a plain inner product, bitwise symmetric in (i, j))
__device__ __forceinline__ void stage_panel(double* dst, int ld, const double* __restrict__ P, int row0, int nrows, const double* sc, int nc) {
  for (int e = threadIdx.x; e < TS * nc; e += blockDim.x) {
    const int r = e / nc, c = e - r * nc;
    const int gr = row0 + r;
    dst[r * ld + c] = (gr < nrows) ? P[(size_t)gr * PM + c] * sc[c] : 0.0;
  }
}
__device__ __forceinline__ void lowrank_tile(double (&F)[4][4], const double* Zi, const double* Zj, int ld, int nc, int ty, int tx) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) F[r][c] = 0.0;
#pragma unroll 4
  for (int a = 0; a < nc; ++a) {
    double zi[4], zj[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) zi[r] = Zi[(ty + 16 * r) * ld + a];
#pragma unroll
    for (int c = 0; c < 4; ++c) zj[c] = Zj[(tx + 16 * c) * ld + a];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) F[r][c] = fma(zi[r], zj[c], F[r][c]);
  }
}
// sqrt(max(theta, 0)) of a block's Ritz values into shared memory; returns the number of leading columns that hold every
// positive Ritz value (they are sorted in descending order by k_rr, so this is their count)
__device__ __forceinline__ int stage_scale(double* sc, const double* th, int p) {
  int nc = 0;
  for (int a = 0; a < p; ++a) if (th[a] > 0.0) nc = a + 1;
  if (threadIdx.x < PM) sc[threadIdx.x] = (threadIdx.x < p && th[threadIdx.x] > 0.0) ? sqrt(th[threadIdx.x]) : 0.0;
  return nc;
}
// shared-memory layout of the Y / U region kernels (doubles): panels of the three blocks (rows i and rows j), the scales,
// a 64 x 17 tile for U; the transposition buffer of store_mirror aliases the panels (they are dead by then)
constexpr int YLD2 = P3 + 1;
constexpr int YOFF_I0 = 0, YOFF_J0 = TS * ZLD, YOFF_I1 = 2 * TS * ZLD, YOFF_J1 = 3 * TS * ZLD, YOFF_I2 = 4 * TS * ZLD,
              YOFF_J2 = 4 * TS * ZLD + TS * YLD2, YOFF_SC = 4 * TS * ZLD + 2 * TS * YLD2, YOFF_UB = YOFF_SC + 3 * PM,
              YOFF_X = YOFF_UB + TS * 17;
// write the transpose of a register tile through shared memory: dst[(col0 + c) * ld + row0 + r] = v[r][c]
__device__ __forceinline__ void store_mirror(double* __restrict__ dst, size_t ld, int row0, int col0, int nrows, int ncols,
                                             const double (&v)[4][4], double* tb, int ty, int tx) {
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) tb[(ty + 16 * r) * (TS + 1) + tx + 16 * c] = v[r][c];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int cc = ty + 16 * r, rr = tx + 16 * c;   // (cc, rr): column / row of the ORIGINAL tile
      if (row0 + rr < nrows && col0 + cc < ncols) dst[(size_t)(col0 + cc) * ld + row0 + rr] = tb[rr * (TS + 1) + cc];
    }
}
__device__ __forceinline__ void lower_tile(int t, int& I, int& J) {
  I = 0;
  while ((I + 1) * (I + 2) / 2 <= t) ++I;
  J = t - I * (I + 1) / 2;
}

// ------------------------------------------------------------------------------------------------------------------
// k_xt: X region (tiles 0 .. tn*tm-1) and Theta region (lower tiles), one fused pass.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) k_xt(BigArgs a) {
  extern __shared__ __align__(16) double sm[];
  const Layout& L = a.L;
  const int slot = a.active[blockIdx.y];
  double* S = node_ptr(a, slot);
  const int n = L.n, m = L.m, N1 = L.N[0];
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double* Zi = sm; double* Zj = Zi + TS * ZLD; double* sc = Zj + TS * ZLD; double* tb = sm;   // (tb aliases the dead panels)
  const double rho = S[L.scal + S_RHO], sig = a.o.sigma, al = a.o.alpha;
  const int nc = stage_scale(sc, S + L.th[0], L.p[0]);
  __syncthreads();
  double* V1 = S + L.V[0];
  const int t = blockIdx.x;
  double F[4][4];
  if (t < L.tn * L.tm) {
    const int I = t / L.tm, J = t - I * L.tm;
    const int i0 = I * TS, j0 = J * TS;
    prefetch_tile(V1 + n, N1, i0, j0, n, m, ty, tx);
    prefetch_tile(S + L.X, m, i0, j0, n, m, ty, tx);
    stage_panel(Zi, ZLD, S + L.Z[0], i0, n, sc, nc);
    stage_panel(Zj, ZLD, S + L.Z[0] + (size_t)n * PM, j0, m, sc, nc);
    __syncthreads();
    lowrank_tile(F, Zi, Zj, ZLD, nc, ty, tx);
    double vnew[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = i0 + ty + 16 * r, j = j0 + tx + 16 * c;
        vnew[r][c] = 0.0;
        if (i < n && j < m) {
          const size_t e = (size_t)i * m + j, ev = (size_t)i * N1 + n + j;
          const double D = V1[ev], x = S[L.X + e], mk = (double)a.Mk[e], am = a.AM[e];
          const double gX = -2.0 * rho * (D - 2.0 * F[r][c]);
          double xt;
          if (!a.sh.on) {
            xt = (sig * x + am + gX) / (mk + sig + 2.0 * rho);
          } else {
            // Shor rows: X = sum_t Xt[t], linear objective -A X on Omega; per-coordinate Sherman-Morrison over the k slices
            //   (sig + 2 rho (cnt + [k>1] cov)) x_t + (2 rho + rho soc) sum_s x_s = sig Xt[t] + A mask + gX + rho GX[t]
            const ShorLayout& SL = a.sh.SL;
            double* Q = a.sh.SS + (size_t)slot * SL.total;
            const unsigned char fl = a.sh.flags[e];
            const int kk = SL.k;
            const double dX = sig + 2.0 * rho * (a.sh.cnt[e] + ((kk > 1 && (fl & 1)) ? 1.0 : 0.0));
            const double cpl = 2.0 * rho + ((fl & 2) ? rho : 0.0);
            double sum = 0.0;
            for (int t = 0; t < kk; ++t) sum += sig * Q[SL.Xt + (size_t)t * SL.C + e] + am + gX + rho * Q[SL.GX + (size_t)t * SL.C + e];
            const double Ssum = sum / (dX + cpl * kk);
            xt = 0.0;
            for (int t = 0; t < kk; ++t) {
              const double rhs = sig * Q[SL.Xt + (size_t)t * SL.C + e] + am + gX + rho * Q[SL.GX + (size_t)t * SL.C + e];
              const double xtt = (rhs - cpl * Ssum) / dX;
              Q[SL.Xtt + (size_t)t * SL.C + e] = xtt;
              xt += xtt;
            }
          }
          const double vn = D + al * (xt - F[r][c]);
          V1[ev] = vn;
          vnew[r][c] = vn;
          S[L.X + e] = al * xt + (1.0 - al) * x;
        }
      }
    store_mirror(V1 + (size_t)n * N1, N1, i0, j0, n, m, vnew, tb, ty, tx);   // V1[n + j][i]
  } else {
    int I, J;
    lower_tile(t - L.tn * L.tm, I, J);
    const int i0 = I * TS, j0 = J * TS;
    prefetch_tile(V1 + (size_t)n * N1 + n, N1, i0, j0, m, m, ty, tx);
    prefetch_tile(S + L.T, m, i0, j0, m, m, ty, tx);
    stage_panel(Zi, ZLD, S + L.Z[0] + (size_t)n * PM, i0, m, sc, nc);
    stage_panel(Zj, ZLD, S + L.Z[0] + (size_t)n * PM, j0, m, sc, nc);
    __syncthreads();
    lowrank_tile(F, Zi, Zj, ZLD, nc, ty, tx);
    double vnew[4][4], tnew[4][4];
    double* T = S + L.T;
    double* V1T = V1 + (size_t)n * N1 + n;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = i0 + ty + 16 * r, j = j0 + tx + 16 * c;
        vnew[r][c] = tnew[r][c] = 0.0;
        if (i < m && j < m) {
          const size_t e = (size_t)i * m + j, ev = (size_t)i * N1 + j;
          const double D = V1T[ev], tt0 = T[e];
          const double gT = -rho * (D - 2.0 * F[r][c]);
          if (a.sh.on && i == j) {       // Shor rows couple Theta~_jj to column j of W: k_shor_col updates this entry
            double* Q = a.sh.SS + (size_t)slot * a.sh.SL.total;
            Q[a.sh.SL.gTd + j] = gT; Q[a.sh.SL.Fd + j] = F[r][c];
            vnew[r][c] = D; tnew[r][c] = tt0;
            continue;
          }
          const double tt = (sig * tt0 - ((i == j) ? a.cT : 0.0) + gT) / (sig + rho);
          const double vn = D + al * (tt - F[r][c]);
          const double tn_ = al * tt + (1.0 - al) * tt0;
          V1T[ev] = vn; T[e] = tn_;
          vnew[r][c] = vn; tnew[r][c] = tn_;
        }
      }
    if (I != J) {
      store_mirror(V1T, N1, i0, j0, m, m, vnew, tb, ty, tx);
      store_mirror(T, m, i0, j0, m, m, tnew, tb, ty, tx);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_y1: Y region (lower tiles) and U region (tn tiles of 64 rows), pass A.  Writes the diagonal-solve values Y~pre, U~pre
// and the per-tile partial sums of the dense rows  R(Y~, U~) = [tr Y~ ; -x_l'U~_j ; -sum_j alpha_lj x_l'U~_j + x_l'Y~x_l].
// rows_part[tile][q]: q = 0 trace, 1 + l k + j -> x_l'U~_j, 1 + L k + l -> x_l'Y~x_l  (signs applied in k_small).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

__global__ void __launch_bounds__(256, 2) k_y1(BigArgs a) {
  extern __shared__ __align__(16) double sm[];
  const Layout& L = a.L;
  const int slot = a.active[blockIdx.y];
  double* S = node_ptr(a, slot);
  const int* NI = node_int(a, slot);
  const int n = L.n, k = L.k, N1 = L.N[0], N2 = L.N[1];
  const int Lc = NI[I_L];
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* Zi_[3] = {sm + YOFF_I0, sm + YOFF_I1, sm + YOFF_I2};
  double* Zj_[3] = {sm + YOFF_J0, sm + YOFF_J1, sm + YOFF_J2};
  const int ld_[3] = {ZLD, ZLD, YLD2};
  int nc_[3];
  double* sc = sm + YOFF_SC;            // [3][PM]
  double* ub = sm + YOFF_UB;            // [TS][17] U tile
  double* tb = sm;                      // [TS*(TS+1)] aliases the panels
  double* xi = sm + YOFF_X;             // [Lc][TS]
  double* xj = xi + (size_t)L.Lcap * TS;
  double* tgs = xj + (size_t)L.Lcap * TS;   // [Lc] tg/rho of the aggregated rows
  double* wred = tgs + L.Lcap + 1;      // [8][rcap]
  const double rho = S[L.scal + S_RHO], sig = a.o.sigma;
  const double dYU = sig + 3.0 * rho;
  const int nYt = L.tn * (L.tn + 1) / 2;
  const int t = blockIdx.x;
  const int rq = 1 + Lc * (k + 1);
  for (int e = threadIdx.x; e < 8 * L.rcap; e += 256) wred[e] = 0.0;
  for (int b = 0; b < 3; ++b) nc_[b] = stage_scale(sc + b * PM, S + L.th[b], L.p[b]);
  for (int l = threadIdx.x; l < Lc; l += 256) tgs[l] = S[L.vg + l] - 2.0 * fmax(S[L.vg + l], 0.0) + S[L.cbe + l];
  __syncthreads();
  double* part = S + L.rows_part + (size_t)t * L.rcap;
  if (t < nYt) {
    int I, J;
    lower_tile(t, I, J);
    const int i0 = I * TS, j0 = J * TS;
    prefetch_tile(S + L.V[0], N1, i0, j0, n, n, ty, tx);
    prefetch_tile(S + L.V[1], N2, i0, j0, n, n, ty, tx);
    prefetch_tile(S + L.V[2], n, i0, j0, n, n, ty, tx);
    prefetch_tile(S + L.Y, n, i0, j0, n, n, ty, tx);
    for (int b = 0; b < 3; ++b) {
      stage_panel(Zi_[b], ld_[b], S + L.Z[b], i0, n, sc + b * PM, nc_[b]);
      stage_panel(Zj_[b], ld_[b], S + L.Z[b], j0, n, sc + b * PM, nc_[b]);
    }
    for (int e = threadIdx.x; e < Lc * TS; e += 256) {
      const int l = e / TS, q = e - l * TS;
      xi[l * TS + q] = (i0 + q < n) ? S[L.xs + (size_t)l * n + i0 + q] : 0.0;
      xj[l * TS + q] = (j0 + q < n) ? S[L.xs + (size_t)l * n + j0 + q] : 0.0;
    }
    __syncthreads();
    double F1[4][4], F2[4][4], F3[4][4], yt[4][4];
    lowrank_tile(F1, Zi_[0], Zj_[0], ld_[0], nc_[0], ty, tx);
    lowrank_tile(F2, Zi_[1], Zj_[1], ld_[1], nc_[1], ty, tx);
    lowrank_tile(F3, Zi_[2], Zj_[2], ld_[2], nc_[2], ty, tx);
    const double v4 = S[L.scal + S_V4];
    const double t4 = v4 - 2.0 * fmax(v4, 0.0) + a.ktr;
    double trp = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int il = ty + 16 * r, jl = tx + 16 * c;
        const int i = i0 + il, j = j0 + jl;
        yt[r][c] = 0.0;
        if (i < n && j < n) {
          const double D1 = S[L.V[0] + (size_t)i * N1 + j], D2 = S[L.V[1] + (size_t)i * N2 + j], D3 = S[L.V[2] + (size_t)i * n + j];
          double g = -(D1 - 2.0 * F1[r][c]) - (D2 - 2.0 * F2[r][c]) + (-D3 - 2.0 * F3[r][c]);
          if (i == j) g += a.a + t4;
          for (int l = 0; l < Lc; ++l) g = fma(tgs[l], xi[l * TS + il] * xj[l * TS + jl], g);
          const double y = (sig * S[L.Y + (size_t)i * n + j] + rho * g) / dYU;
          yt[r][c] = y;
          S[L.Yt + (size_t)i * n + j] = y;
          if (i == j) trp += y;
        }
      }
    if (I != J) store_mirror(S + L.Yt, n, i0, j0, n, n, yt, tb, ty, tx);
    // dense-row partial sums over this tile (off-diagonal tiles count twice in x'Yx)
    const double wgt = (I != J) ? 2.0 : 1.0;
    trp = warp_sum(trp);
    if (lane == 0) wred[warp * L.rcap + 0] = trp;
    for (int l = 0; l < Lc; ++l) {
      double q = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) q = fma(xi[l * TS + ty + 16 * r] * xj[l * TS + tx + 16 * c], yt[r][c], q);
      q = warp_sum(q);
      if (lane == 0) wred[warp * L.rcap + 1 + Lc * k + l] = wgt * q;
    }
  } else {
    // U tile: rows i0 .. i0+63, columns 0..k-1.  U~pre = (sig U + rho gU) / dYU,
    // gU/rho = -2 (D2u - 2 F2u) - (v5 - 2 s5) - sum_l x_li (tv_lj + tg_l alpha_lj)
    const int i0 = (t - nYt) * TS;
    double* Zi = Zi_[1]; double* Zj = Zj_[1];
    const int nc1 = nc_[1];
    stage_panel(Zi, ZLD, S + L.Z[1], i0, n, sc + PM, nc1);
    // the k rows n .. n+k-1 of Z2 (scaled)
    for (int e = threadIdx.x; e < k * PM; e += 256) {
      const int j = e / PM, c = e - j * PM;
      Zj[j * ZLD + c] = S[L.Z[1] + (size_t)(n + j) * PM + c] * sc[PM + c];
    }
    for (int e = threadIdx.x; e < Lc * TS; e += 256) {
      const int l = e / TS, q = e - l * TS;
      xi[l * TS + q] = (i0 + q < n) ? S[L.xs + (size_t)l * n + i0 + q] : 0.0;
    }
    __syncthreads();
    const double sa = a.sa;
    for (int e = threadIdx.x; e < TS * k; e += 256) {
      const int il = e / k, j = e - il * k;
      const int i = i0 + il;
      double ut = 0.0;
      if (i < n) {
        double F = 0.0;
        for (int c = 0; c < nc1; ++c) F = fma(Zi[il * ZLD + c], Zj[j * ZLD + c], F);
        const double D = S[L.V[1] + (size_t)i * N2 + n + j];
        const double v5 = S[L.v5 + (size_t)i * k + j];
        const double lo5 = (i >= n - k + j) ? 0.0 : -sa;
        const double s5 = clampd(v5, lo5, sa);
        double g = -2.0 * (D - 2.0 * F) - (v5 - 2.0 * s5);
        for (int l = 0; l < Lc; ++l) {
          const double vv = S[L.vv + (size_t)l * k + j];
          const double sv = clampd(vv, S[L.clb + (size_t)l * k + j], S[L.cub + (size_t)l * k + j]);
          g -= xi[l * TS + il] * ((vv - 2.0 * sv) + tgs[l] * S[L.cal + (size_t)l * k + j]);
        }
        ut = (sig * S[L.U + (size_t)i * k + j] + rho * g) / dYU;
        S[L.Ut + (size_t)i * k + j] = ut;
      }
      ub[il * 17 + j] = ut;           // keep the tile for the row sums below
    }
    __syncthreads();
    // x_l'U~_j over the 64 rows of this tile: one (l, j) per thread
    for (int e = threadIdx.x; e < Lc * k; e += 256) {
      const int l = e / k, j = e - l * k;
      double q = 0.0;
      for (int il = 0; il < TS; ++il) q = fma(xi[l * TS + il], ub[il * 17 + j], q);
      part[1 + l * k + j] = q;
    }
    if (threadIdx.x == 0) part[0] = 0.0;
    for (int l = threadIdx.x; l < Lc; l += 256) part[1 + Lc * k + l] = 0.0;
    return;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < rq; q += 256) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += wred[w * L.rcap + q];
    part[q] = s;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_small: one CTA per node.  R0 = dense rows of (Y~pre, U~pre); cw = Minv R0; Rf = R0 - G cw = rows of the corrected
// (Y~, U~); v-updates of the scalar rows (trace row, cut rows).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_small(BigArgs a) {
  extern __shared__ __align__(16) double sm[];
  const Layout& L = a.L;
  const int slot = a.active[blockIdx.x];
  double* S = node_ptr(a, slot);
  const int* NI = node_int(a, slot);
  const int k = L.k, Lc = NI[I_L];
  const int rq = 1 + Lc * (k + 1);
  double* Rw = sm; double* R0 = Rw + L.rcap; double* cw = R0 + L.rcap; double* Rf = cw + L.rcap;
  for (int q = threadIdx.x; q < rq; q += 128) {
    double s = 0.0;
    for (int t = 0; t < L.tilesYU; ++t) s += S[L.rows_part + (size_t)t * L.rcap + q];
    Rw[q] = s;
  }
  __syncthreads();
  // R(Y~, U~): [tr ; -x'U ; -sum_j alpha x'U + x'Yx]
  for (int q = threadIdx.x; q < rq; q += 128) {
    double v;
    if (q == 0) v = Rw[0];
    else if (q < 1 + Lc * k) v = -Rw[q];
    else {
      const int l = q - 1 - Lc * k;
      v = Rw[q];
      for (int j = 0; j < k; ++j) v -= S[L.cal + (size_t)l * k + j] * Rw[1 + l * k + j];
    }
    R0[q] = v;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < rq; q += 128) {
    double s = 0.0;
    for (int c = 0; c < rq; ++c) s = fma(S[L.Minv + (size_t)q * L.rcap + c], R0[c], s);
    cw[q] = s;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < rq; q += 128) {
    double s = R0[q];
    for (int c = 0; c < rq; ++c) s -= S[L.G + (size_t)q * L.rcap + c] * cw[c];
    Rf[q] = s;
    S[L.cw + q] = cw[q];
  }
  __syncthreads();
  const double al = a.o.alpha;
  if (threadIdx.x == 0) {
    const double v4 = S[L.scal + S_V4];
    const double z4 = a.ktr - Rf[0];
    S[L.scal + S_V4] = v4 + al * (z4 - fmax(v4, 0.0));
  }
  for (int e = threadIdx.x; e < Lc * k; e += 128) {
    const double vv = S[L.vv + e];
    const double zv = -Rf[1 + e];
    S[L.vv + e] = vv + al * (zv - clampd(vv, S[L.clb + e], S[L.cub + e]));
  }
  for (int l = threadIdx.x; l < Lc; l += 128) {
    const double vg = S[L.vg + l];
    const double zg = S[L.cbe + l] - Rf[1 + Lc * k + l];
    S[L.vg + l] = vg + al * (zg - fmax(vg, 0.0));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_y2: Y and U regions, pass B.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_y2(BigArgs a) {
  extern __shared__ __align__(16) double sm[];
  const Layout& L = a.L;
  const int slot = a.active[blockIdx.y];
  double* S = node_ptr(a, slot);
  const int* NI = node_int(a, slot);
  const int n = L.n, k = L.k, N1 = L.N[0], N2 = L.N[1];
  const int Lc = NI[I_L];
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double* Zi_[3] = {sm + YOFF_I0, sm + YOFF_I1, sm + YOFF_I2};
  double* Zj_[3] = {sm + YOFF_J0, sm + YOFF_J1, sm + YOFF_J2};
  const int ld_[3] = {ZLD, ZLD, YLD2};
  int nc_[3];
  double* sc = sm + YOFF_SC;
  double* tb = sm;                          // aliases the panels
  double* xi = sm + YOFF_X;
  double* xj = xi + (size_t)L.Lcap * TS;
  double* cg = xj + (size_t)L.Lcap * TS;    // [Lc] cw of the aggregated rows
  const double al = a.o.alpha;
  const int nYt = L.tn * (L.tn + 1) / 2;
  const int t = blockIdx.x;
  for (int b = 0; b < 3; ++b) nc_[b] = stage_scale(sc + b * PM, S + L.th[b], L.p[b]);
  for (int l = threadIdx.x; l < Lc; l += 256) cg[l] = S[L.cw + 1 + Lc * k + l];
  __syncthreads();
  const double cw0 = S[L.cw + 0];
  if (t < nYt) {
    int I, J;
    lower_tile(t, I, J);
    const int i0 = I * TS, j0 = J * TS;
    prefetch_tile(S + L.Yt, n, i0, j0, n, n, ty, tx);
    prefetch_tile(S + L.V[0], N1, i0, j0, n, n, ty, tx);
    prefetch_tile(S + L.V[1], N2, i0, j0, n, n, ty, tx);
    prefetch_tile(S + L.V[2], n, i0, j0, n, n, ty, tx);
    prefetch_tile(S + L.Y, n, i0, j0, n, n, ty, tx);
    for (int b = 0; b < 3; ++b) {
      stage_panel(Zi_[b], ld_[b], S + L.Z[b], i0, n, sc + b * PM, nc_[b]);
      stage_panel(Zj_[b], ld_[b], S + L.Z[b], j0, n, sc + b * PM, nc_[b]);
    }
    for (int e = threadIdx.x; e < Lc * TS; e += 256) {
      const int l = e / TS, q = e - l * TS;
      xi[l * TS + q] = (i0 + q < n) ? S[L.xs + (size_t)l * n + i0 + q] : 0.0;
      xj[l * TS + q] = (j0 + q < n) ? S[L.xs + (size_t)l * n + j0 + q] : 0.0;
    }
    __syncthreads();
    double F1[4][4], F2[4][4], F3[4][4], o1[4][4], o2[4][4], o3[4][4], oy[4][4];
    lowrank_tile(F1, Zi_[0], Zj_[0], ld_[0], nc_[0], ty, tx);
    lowrank_tile(F2, Zi_[1], Zj_[1], ld_[1], nc_[1], ty, tx);
    lowrank_tile(F3, Zi_[2], Zj_[2], ld_[2], nc_[2], ty, tx);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int il = ty + 16 * r, jl = tx + 16 * c;
        const int i = i0 + il, j = j0 + jl;
        o1[r][c] = o2[r][c] = o3[r][c] = oy[r][c] = 0.0;
        if (i < n && j < n) {
          double cy = (i == j) ? cw0 : 0.0;
          for (int l = 0; l < Lc; ++l) cy = fma(cg[l], xi[l * TS + il] * xj[l * TS + jl], cy);
          const double yt = S[L.Yt + (size_t)i * n + j] - cy;
          const size_t e1 = L.V[0] + (size_t)i * N1 + j, e2 = L.V[1] + (size_t)i * N2 + j, e3 = L.V[2] + (size_t)i * n + j;
          const double v1 = S[e1] + al * (yt - F1[r][c]);
          const double v2 = S[e2] + al * (yt - F2[r][c]);
          const double d3 = S[e3];
          const double v3 = d3 + al * (((i == j) ? a.a : 0.0) - yt - d3 - F3[r][c]);
          const size_t ey = L.Y + (size_t)i * n + j;
          const double yn = al * yt + (1.0 - al) * S[ey];
          S[e1] = v1; S[e2] = v2; S[e3] = v3; S[ey] = yn;
          o1[r][c] = v1; o2[r][c] = v2; o3[r][c] = v3; oy[r][c] = yn;
        }
      }
    if (I != J) {
      store_mirror(S + L.V[0], N1, i0, j0, n, n, o1, tb, ty, tx);
      store_mirror(S + L.V[1], N2, i0, j0, n, n, o2, tb, ty, tx);
      store_mirror(S + L.V[2], n, i0, j0, n, n, o3, tb, ty, tx);
      store_mirror(S + L.Y, n, i0, j0, n, n, oy, tb, ty, tx);
    }
  } else {
    const int i0 = (t - nYt) * TS;
    double* Zi = Zi_[1]; double* Zj = Zj_[1];
    const int nc1 = nc_[1];
    stage_panel(Zi, ZLD, S + L.Z[1], i0, n, sc + PM, nc1);
    for (int e = threadIdx.x; e < k * PM; e += 256) {
      const int j = e / PM, c = e - j * PM;
      Zj[j * ZLD + c] = S[L.Z[1] + (size_t)(n + j) * PM + c] * sc[PM + c];
    }
    for (int e = threadIdx.x; e < Lc * TS; e += 256) {
      const int l = e / TS, q = e - l * TS;
      xi[l * TS + q] = (i0 + q < n) ? S[L.xs + (size_t)l * n + i0 + q] : 0.0;
    }
    __syncthreads();
    const double sa = a.sa;
    for (int e = threadIdx.x; e < TS * k; e += 256) {
      const int il = e / k, j = e - il * k;
      const int i = i0 + il;
      if (i >= n) continue;
      double F = 0.0;
      for (int c = 0; c < nc1; ++c) F = fma(Zi[il * ZLD + c], Zj[j * ZLD + c], F);
      // cU_ij = - sum_l x_li (cv_lj + cg_l alpha_lj);  U~ = U~pre - cU
      double cu = 0.0;
      for (int l = 0; l < Lc; ++l) cu -= xi[l * TS + il] * (S[L.cw + 1 + l * k + j] + cg[l] * S[L.cal + (size_t)l * k + j]);
      const double ut = S[L.Ut + (size_t)i * k + j] - cu;
      const size_t e2 = L.V[1] + (size_t)i * N2 + n + j, e2t = L.V[1] + (size_t)(n + j) * N2 + i;
      const double v2 = S[e2] + al * (ut - F);
      S[e2] = v2; S[e2t] = v2;
      const double v5 = S[L.v5 + (size_t)i * k + j];
      const double lo5 = (i >= n - k + j) ? 0.0 : -sa;
      S[L.v5 + (size_t)i * k + j] = v5 + al * (ut - clampd(v5, lo5, sa));
      const size_t eu = L.U + (size_t)i * k + j;
      S[eu] = al * ut + (1.0 - al) * S[eu];
    }
    if (t == nYt) {
      // the k x k identity corner of block 2: z = delta
      for (int e = threadIdx.x; e < k * k; e += 256) {
        const int i = e / k, j = e - i * k;
        double F = 0.0;
        for (int c = 0; c < nc1; ++c) F = fma(Zj[i * ZLD + c], Zj[j * ZLD + c], F);
        const size_t ev = L.V[1] + (size_t)(n + i) * N2 + n + j;
        S[ev] = S[ev] + al * (((i == j) ? 1.0 : 0.0) - F);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_prod: Wout = side * V_b P, P a PM-column panel (Z for which = 0, R for which = 1), FP64 DMMA m8n8k4.  One CTA = 64 rows
// of one (node, block): V streams through shared memory in 64 x 32 chunks (register-prefetched).  Epilogue: partial Gram
// matrices of this row tile, partA = Z_tile' Wout_tile and (which = 1) partB = R_tile' Wout_tile.
// ------------------------------------------------------------------------------------------------------------------
constexpr int KC = 32, VLD = 36, PLD = PM + 4;   // 36 = 4 mod 16: the 8 x 4 (row / k, column) fragment loads of a half-warp fall in 16 distinct bank pairs
constexpr int NTL = PM / 8;     // DMMA n-tiles per warp

__global__ void __launch_bounds__(128) k_prod(BigArgs a, int which) {
  __shared__ __align__(16) double Vs[TS * VLD];      // V chunk; the W tile (ld ZLD) aliases it in the epilogue
  __shared__ __align__(16) double Ps[KC * PLD];
  __shared__ double Ls[TS * ZLD];
  const Layout& L = a.L;
  const int b = blockIdx.z;
  if ((int)blockIdx.x >= L.nt[b]) return;
  const int slot = a.active[blockIdx.y];
  const int* NI = node_int(a, slot);
  if (a.step > 0 && !NI[I_MORE + b]) return;
  double* S = node_ptr(a, slot);
  const int N = L.N[b];
  const double sd = (b == 2) ? -1.0 : 1.0;
  const double* __restrict__ V = S + L.V[b];
  const double* __restrict__ P = S + (which ? L.R[b] : L.Z[b]);
  double* Wout = S + (which ? L.W2[b] : L.W[b]);
  const int r0 = blockIdx.x * TS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  // live 8-column tiles of the panel: Z has p columns (16 for the third block); R only the window k_resid kept (the
  // residuals of the positive Ritz pairs, `window` guard columns and the probe) -- the other columns are exact zeros
  unsigned ntmask = 0;
  {
    const int p = L.p[b];
    int na = p;
    if (which && a.o.window > 0 && a.it >= 2) {
      int rprev = 0;
      for (int cc = 0; cc < p; ++cc) rprev += (S[L.th[b] + cc] > 0.0) ? 1 : 0;
      na = min(p, rprev + a.o.window);
    }
    for (int nt = 0; nt < NTL; ++nt)
      if (nt * 8 < na || (which && p > 1 && nt == (p - 1) / 8)) ntmask |= 1u << nt;
  }
  double c[2][NTL][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NTL; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  double pv[16], pp[KC * PM / 128];
  auto fetch = [&](int kc) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int idx = q * 128 + tid, row = idx >> 5, col = idx & 31;
      const int gr = r0 + row, gc = kc + col;
      pv[q] = (gr < N && gc < N) ? V[(size_t)gr * N + gc] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < KC * PM / 128; ++q) {
      const int idx = q * 128 + tid, row = idx / PM, col = idx % PM;
      pp[q] = (kc + row < N) ? P[(size_t)(kc + row) * PM + col] : 0.0;
    }
  };
  fetch(0);
  for (int kc = 0; kc < N; kc += KC) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int idx = q * 128 + tid, row = idx >> 5, col = idx & 31;
      Vs[row * VLD + col] = pv[q];
    }
#pragma unroll
    for (int q = 0; q < KC * PM / 128; ++q) {
      const int idx = q * 128 + tid, row = idx / PM, col = idx % PM;
      Ps[row * PLD + col] = pp[q];
    }
    __syncthreads();
    if (kc + KC < N) fetch(kc + KC);
    const int rb = warp * 16;
#pragma unroll
    for (int kk = 0; kk < KC / 4; ++kk) {
      const double a0 = Vs[(rb + g) * VLD + kk * 4 + t4], a1 = Vs[(rb + 8 + g) * VLD + kk * 4 + t4];
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        if ((ntmask >> nt) & 1u) {
          const double bb = Ps[(kk * 4 + t4) * PLD + nt * 8 + g];
          dmma884(c[0][nt][0], c[0][nt][1], a0, bb, c[0][nt][0], c[0][nt][1]);
          dmma884(c[1][nt][0], c[1][nt][1], a1, bb, c[1][nt][0], c[1][nt][1]);
        }
      }
    }
  }
  __syncthreads();
  double* Ws = Vs;                 // 64 x ZLD <= 64 x VLD
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
      const int row = warp * 16 + mt * 8 + g, col = nt * 8 + 2 * t4;
      Ws[row * ZLD + col] = sd * c[mt][nt][0];
      Ws[row * ZLD + col + 1] = sd * c[mt][nt][1];
    }
  __syncthreads();
  for (int e = tid; e < TS * PM; e += 128) {
    const int row = e / PM, col = e % PM;
    if (r0 + row < N) Wout[(size_t)(r0 + row) * PM + col] = Ws[row * ZLD + col];
    else Ws[row * ZLD + col] = 0.0;
  }
  for (int pass = 0; pass <= which; ++pass) {
    const double* Lp = S + (pass ? L.R[b] : L.Z[b]);
    __syncthreads();
    for (int e = tid; e < TS * PM; e += 128) {
      const int row = e / PM, col = e % PM;
      Ls[row * ZLD + col] = (r0 + row < N) ? Lp[(size_t)(r0 + row) * PM + col] : 0.0;
    }
    __syncthreads();
    double* part = S + (pass ? L.partB[b] : L.partA[b]) + (size_t)blockIdx.x * PM * PM;
    // 2 x 4 outputs per thread (16 x 8 threads cover 32 x 32)
    const int ai = (tid >> 3) * 2, bi = (tid & 7) * 4;
    double sacc[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) sacc[u][v] = 0.0;
#pragma unroll 4
    for (int r = 0; r < TS; ++r) {
      const double l0 = Ls[r * ZLD + ai], l1 = Ls[r * ZLD + ai + 1];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const double w = Ws[r * ZLD + bi + v];
        sacc[0][v] = fma(l0, w, sacc[0][v]); sacc[1][v] = fma(l1, w, sacc[1][v]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) part[(ai + u) * PM + bi + v] = sacc[u][v];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_resid, pass 0: R = W - Z H (H = sum of partA, or diag(theta) after the first step of an iteration), probe column;
//                  partB = Z_tile' R_tile.
//          pass 1: C = sum partB; R -= Z C; partA = Z_tile' R_tile.        (two explicit projections against Z: "twice is
//          pass 2: C = sum partA; R -= Z C; partB = R_tile' R_tile.         enough", see oracle/bigblock.py)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resid(BigArgs a, int pass) {
  __shared__ double Hs[PM * ZLD];
  __shared__ double Zs[TS * ZLD];
  __shared__ double Rs[TS * ZLD];
  __shared__ double red[32];
  const Layout& L = a.L;
  const int b = blockIdx.z;
  if ((int)blockIdx.x >= L.nt[b]) return;
  const int slot = a.active[blockIdx.y];
  const int* NI = node_int(a, slot);
  if (a.step > 0 && !NI[I_MORE + b]) return;
  double* S = node_ptr(a, slot);
  const int N = L.N[b], tid = threadIdx.x;
  const int r0 = blockIdx.x * TS;
  double h2 = 0.0;
  for (int e = tid; e < PM * PM; e += 256) {
    const int ai = e / PM, bi = e % PM;
    double s = 0.0;
    if (pass == 0 && a.step > 0) {
      s = (ai == bi && ai < L.p[b]) ? S[L.th[b] + ai] : 0.0;   // (dead columns carry a -1e300 sentinel)
    } else {
      const double* part = S + ((pass & 1) ? L.partB[b] : L.partA[b]);
      for (int t = 0; t < L.nt[b]; ++t) s += part[(size_t)t * PM * PM + e];
    }
    Hs[ai * ZLD + bi] = s;
    h2 += s * s;
    if (pass == 0 && blockIdx.x == 0) S[L.H[b] + e] = s;
  }
  const double* src = S + (pass ? L.R[b] : L.W[b]);
  for (int e = tid; e < TS * PM; e += 256) {
    const int row = e / PM, col = e % PM;
    const bool ok = r0 + row < N;
    Zs[row * ZLD + col] = ok ? S[L.Z[b] + (size_t)(r0 + row) * PM + col] : 0.0;
    Rs[row * ZLD + col] = ok ? src[(size_t)(r0 + row) * PM + col] : 0.0;
  }
  __syncthreads();
  // pass 0: the last panel column carries a fresh pseudo-random probe instead of its residual, so that the trial space
  // [Z, R] stays generic: an eigenvector exactly orthogonal to Z and to every residual (e.g. the identity corner of
  // [Y U; U' I] at a node without cuts, where U = 0) would otherwise never be found again once it has left the panel
  double amp = 0.0;
  if (pass == 0) {
    h2 = block_sum(h2, red);
    amp = 1e-3 * sqrt(h2 / (double)N);
  }
  {
    constexpr int CPT = TS * PM / 256;      // outputs per thread (8): one row, CPT consecutive columns
    const int row = tid / (PM / CPT), cb = (tid % (PM / CPT)) * CPT;
    double acc[CPT];
#pragma unroll
    for (int q = 0; q < CPT; ++q) acc[q] = Rs[row * ZLD + cb + q];
#pragma unroll 4
    for (int c = 0; c < PM; ++c) {
      const double z = Zs[row * ZLD + c];
#pragma unroll
      for (int q = 0; q < CPT; ++q) acc[q] = fma(-z, Hs[c * ZLD + cb + q], acc[q]);
    }
    if (pass == 0) {
      const int pc = L.p[b] - 1;
      if (pc > 0) {
#pragma unroll
        for (int q = 0; q < CPT; ++q)
          if (cb + q == pc)
            acc[q] = amp * hash_unit((unsigned long long)(r0 + row), (unsigned long long)(a.it * 64 + a.step * 4 + b), 12345ull);
      }
      // residual window: only the residuals of the positive Ritz pairs and of the first `window` guard columns (and the
      // probe) enter the trial space; the Rayleigh-Ritz still runs over ALL of Z, so no Ritz value ever gets worse.  The
      // zero columns are dropped by the Cholesky guard and k_rr's Jacobi runs on p + (r + window + 1) indices instead of 2 p.
      if (a.o.window > 0 && a.it >= 2) {
        int rprev = 0;
        for (int c = 0; c < L.p[b]; ++c) rprev += (S[L.th[b] + c] > 0.0) ? 1 : 0;
        const int na = rprev + a.o.window;
#pragma unroll
        for (int q = 0; q < CPT; ++q)
          if (cb + q >= na && cb + q != pc) acc[q] = 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
      Rs[row * ZLD + cb + q] = acc[q];
      if (r0 + row < N) S[L.R[b] + (size_t)(r0 + row) * PM + cb + q] = acc[q];
    }
  }
  __syncthreads();
  {
    const double* Lm = (pass == 2) ? Rs : Zs;
    double* part = S + ((pass & 1) ? L.partA[b] : L.partB[b]) + (size_t)blockIdx.x * PM * PM;
    // 2 x 2 outputs per thread (16 x 16 threads cover PM x PM = 32 x 32): one shared-memory load per FMA instead of two
    static_assert(PM == 32, "Gram tiling assumes PM = 32");
    const int ai = (tid >> 4) * 2, bi = (tid & 15) * 2;
    double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
#pragma unroll 8
    for (int r = 0; r < TS; ++r) {
      const double l0 = Lm[r * ZLD + ai], l1 = Lm[r * ZLD + ai + 1], r0v = Rs[r * ZLD + bi], r1v = Rs[r * ZLD + bi + 1];
      s00 = fma(l0, r0v, s00); s01 = fma(l0, r1v, s01); s10 = fma(l1, r0v, s10); s11 = fma(l1, r1v, s11);
    }
    part[ai * PM + bi] = s00; part[ai * PM + bi + 1] = s01; part[(ai + 1) * PM + bi] = s10; part[(ai + 1) * PM + bi + 1] = s11;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// guarded Cholesky-QR factor of the PM x PM Gram matrix M (ld ZLD) by one warp.  Columns are equilibrated first
// (M~ = D^-1 M D^-1, D = sqrt(diag M)) so that the guards do not depend on the relative scale of the columns: a column is
// dropped when its squared norm is <= floor2, or <= rel_small * the largest one (the probe column is exempt and does not
// count for the largest), or when its pivot in M~ is <= piv_rel (nearly dependent on earlier columns).
// Tm = D^-1 L~^-T with zero columns for dropped ones: P Tm is orthonormal on the kept columns.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void chol_guard_warp(double* M, double* Lm, double* Tm, int* valid, int lane, double piv_rel, double floor2,
                                                int pcols, double rel_small = 0.0, int probe = -1) {
  __shared__ double dscale[PM];
  {
    const double d = (lane < PM) ? M[lane * ZLD + lane] : 0.0;
    const double dmax = fmax(warp_max((lane < pcols && lane != probe) ? d : 0.0), 0.0);
    if (lane < PM) {
      const bool ok = (lane < pcols) && (d > floor2) && (d > 0.0) && (d > rel_small * dmax || lane == probe);
      dscale[lane] = ok ? 1.0 / sqrt(d) : 0.0;
    }
  }
  for (int e = lane; e < PM * ZLD; e += 32) { Lm[e] = 0.0; Tm[e] = 0.0; }
  __syncwarp();
  for (int e = lane; e < PM * PM; e += 32) {
    const int i = e / PM, j = e % PM;
    M[i * ZLD + j] *= dscale[i] * dscale[j];
  }
  __syncwarp();
  for (int j = 0; j < PM; ++j) {
    double v = M[j * ZLD + j];
    for (int c = 0; c < j; ++c) v -= Lm[j * ZLD + c] * Lm[j * ZLD + c];
    const bool ok = (dscale[j] > 0.0) && (v > piv_rel);
    if (lane == 0) valid[j] = ok ? 1 : 0;
    if (ok) {
      const double d = sqrt(v);
      if (lane > j && lane < PM) {
        double sacc = M[lane * ZLD + j];
        for (int c = 0; c < j; ++c) sacc -= Lm[lane * ZLD + c] * Lm[j * ZLD + c];
        Lm[lane * ZLD + j] = sacc / d;
      }
      if (lane == j) Lm[j * ZLD + j] = d;
    } else {
      if (lane == j) Lm[j * ZLD + j] = 1.0;       // dropped column: unit pivot, zero column
    }
    __syncwarp();
  }
  for (int j = 0; j < PM; ++j) {
    if (!valid[j]) for (int c = lane; c < j; c += 32) Lm[j * ZLD + c] = 0.0;
  }
  __syncwarp();
  // L~inv by forward substitution, one column per lane, built in place in Tm: Tm[c][i] = L~inv[i][c] (then scaled)
  if (lane < PM) {
    const int cidx = lane;
    for (int i = cidx; i < PM; ++i) {
      double sacc = (i == cidx) ? 1.0 : 0.0;
      for (int c = cidx; c < i; ++c) sacc -= Lm[i * ZLD + c] * Tm[cidx * ZLD + c];
      Tm[cidx * ZLD + i] = sacc / Lm[i * ZLD + i];
    }
    for (int i = 0; i < PM; ++i) Tm[cidx * ZLD + i] = (i >= cidx && valid[i] && valid[cidx]) ? dscale[cidx] * Tm[cidx * ZLD + i] : 0.0;
  }
  __syncwarp();
}

// k_gram: M = sum over tiles of partB (R'R, written by k_resid pass 2) saved to L.Q[b] before k_prod(which = 1) reuses it.
__global__ void __launch_bounds__(256) k_gram(BigArgs a) {
  const Layout& L = a.L;
  const int b = blockIdx.y;
  const int slot = a.active[blockIdx.x];
  const int* NI = node_int(a, slot);
  if (a.step > 0 && !NI[I_MORE + b]) return;
  double* S = node_ptr(a, slot);
  for (int e = threadIdx.x; e < PM * PM; e += 256) {
    double s = 0.0;
    for (int t = 0; t < L.nt[b]; ++t) s += S[L.partB[b] + (size_t)t * PM * PM + e];
    S[L.Q[b] + e] = s;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-parallel cyclic Jacobi (256 threads) on the n2 x n2 symmetric matrix Ac (shared memory, ld JLD, n2 even <= JN); Qc
// accumulates the rotations (identity on entry).  Round-robin ordering: n2 - 1 rounds of n2 / 2 disjoint pairs; per round the
// first n2 / 2 threads compute the rotations, then all row updates, then all column updates of Ac and Qc (all loads of a phase
// are issued before its first store: the pairs are disjoint).  The thread -> (pair, column) map is division-free.
// ------------------------------------------------------------------------------------------------------------------
constexpr int JN = 2 * PM, JLD = JN + 1;
__device__ __forceinline__ void jacobi_cta(double* Ac, double* Qc, int n2, int max_sweeps, double off_tol, double amax, double* cs, int* pq,
                                           int* flag, double* red) {
  // Cyclic Jacobi, round-robin ordering, n2 / 2 disjoint rotations per round.  Two barriers per round: the rotation parameters
  // (one warp), then ONE pass that applies J' A J block-wise -- the 2 x 2 block (pair i, pair j) needs only its own four
  // entries for the row rotation of pair i followed by the column rotation of pair j -- and the column rotations of Q.
  // flag[3..5] rotate over the rounds (a round whose rotations are all identities is skipped without a barrier).
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int npair = n2 >> 1;
  int* rf = flag + 3;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    __syncthreads();
    if (tid < 3) rf[tid] = 0;
    double off = 0.0;
    {
      const int cj = tid & (JN - 1);
      if (cj < n2)
        for (int i = tid >> 6; i < n2; i += 4)
          if (i != cj) off = fmax(off, fabs(Ac[i * JLD + cj]));
    }
    off = block_max(off, red);
    if (off <= off_tol || off < 1e-300) break;
    for (int rnd = 0; rnd < n2 - 1; ++rnd) {
      const int fcur = rnd % 3;
      if (tid < npair) {
        int p_, q_;
        if (tid == 0) { p_ = n2 - 1; q_ = rnd; rf[(rnd + 1) % 3] = 0; }
        else { p_ = (rnd + tid) % (n2 - 1); q_ = (rnd + n2 - 1 - tid) % (n2 - 1); }
        if (p_ > q_) { const int t_ = p_; p_ = q_; q_ = t_; }
        const double apq = Ac[p_ * JLD + q_], app = Ac[p_ * JLD + p_], aqq = Ac[q_ * JLD + q_];
        double c_ = 1.0, s_ = 0.0;
        if (fabs(apq) > 1e-16 * amax && fabs(apq) > 1e-17 * (fabs(app) + fabs(aqq))) {
          const double tau = (aqq - app) / (2.0 * apq);
          const double t_ = ((tau >= 0.0) ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c_ = rsqrt(1.0 + t_ * t_);
          s_ = t_ * c_;
          rf[fcur] = 1;
        }
        cs[tid * 2 + 0] = c_; cs[tid * 2 + 1] = s_;
        pq[tid * 2 + 0] = p_; pq[tid * 2 + 1] = q_;
      }
      __syncthreads();
      if (!rf[fcur]) continue;
      // lane = pair j (its rotation stays in registers for the whole pass), warp + 8 u = pair i: no index arithmetic beyond
      // two row offsets per block; consecutive lanes hold consecutive columns (round-robin ordering), so the shared-memory
      // accesses of a warp fall in distinct banks except at the wrap-around
      if (lane < npair) {
        const int pj = pq[2 * lane], qj = pq[2 * lane + 1];
        const double cj = cs[2 * lane], sj = cs[2 * lane + 1];
#pragma unroll
        for (int u = 0; u < PM / 8; ++u) {
          const int i = wrp + 8 * u;
          if (i < npair) {
            double* rp = Ac + pq[2 * i] * JLD;
            double* rq = Ac + pq[2 * i + 1] * JLD;
            const double ci = cs[2 * i], si = cs[2 * i + 1];
            const double app = rp[pj], apq = rp[qj], aqp = rq[pj], aqq = rq[qj];
            const double rpp = ci * app - si * aqp, rqp = si * app + ci * aqp;
            const double rpq = ci * apq - si * aqq, rqq = si * apq + ci * aqq;
            rp[pj] = cj * rpp - sj * rpq; rp[qj] = sj * rpp + cj * rpq;
            rq[pj] = cj * rqp - sj * rqq; rq[qj] = sj * rqp + cj * rqq;
          }
        }
        for (int row = wrp; row < n2; row += 8) {
          double* qr = Qc + row * JLD;
          const double qp = qr[pj], qq = qr[qj];
          qr[pj] = cj * qp - sj * qq;
          qr[qj] = sj * qp + cj * qq;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// k_rr: one CTA (256 threads) per (node, block).  M = R'R (saved by k_gram) -> guarded Cholesky -> T = D^-1 L~^-T; the second
// product gave partA = Z'(V R), partB = R'(V R): Xc = (Z'VR) T, C = T'(R'VR) T; Rayleigh-Ritz on [[H, Xc], [Xc', C]]
// (2 PM x 2 PM) by a CTA-parallel cyclic Jacobi (round-robin pairs; per round: PM rotations, then all row updates, then all
// column updates of A and of the eigenvector accumulator Q); keeps the p largest Ritz pairs; writes Q' (2 PM x PM, the bottom
// half already multiplied by T so that Znew = Z Qtop + R Qbot), theta, and the step control flags.
// The rotations keep Q orthogonal whatever the number of sweeps; the projected matrix is nearly diagonal once the tracker has
// locked on (off-diagonal = Ritz residual), so `jacobi_sweeps` = 3 resolves it to rounding there, and an unconverged
// start-phase Rayleigh-Ritz is finished by the following tracker steps.
// ------------------------------------------------------------------------------------------------------------------
// shared memory of k_rr: the two 2p x 2p matrices only (69 KB: three CTAs per SM).  The p x p work matrices of the set-up (M, the
// Cholesky factor, T = L~^-1, X0, C0, C0 T) live inside the second matrix before the Jacobi needs it, T is parked in the
// bottom half of the node's Q' output during the Jacobi and comes back for the last product.
constexpr int RR_X0 = 0, RR_C0 = PM * PM, RR_TM = 2 * PM * PM, RR_MM = 2 * PM * PM + PM * ZLD;   // offsets inside the Q region
static_assert(RR_MM + PM * ZLD <= JN * JLD, "k_rr set-up matrices must fit the second Jacobi matrix");
constexpr size_t RR_SMEM = ((size_t)2 * JN * JLD + 2 * PM + JN + 64) * sizeof(double) + (3 * PM + 2 * JN + 8) * sizeof(int);

__global__ void __launch_bounds__(256, 3) k_rr(BigArgs a) {
  extern __shared__ __align__(16) double sm[];
  double* A = sm;                       // [JN][JLD]
  double* Q = A + JN * JLD;             // [JN][JLD]   (set-up matrices before the Jacobi, see RR_*)
  double* cs = Q + JN * JLD;            // [PM][2]
  double* lam = cs + 2 * PM;            // [JN]
  double* red = lam + JN;               // [64]
  int* valid = reinterpret_cast<int*>(red + 64);   // [PM]
  int* pq = valid + PM;                 // [PM][2]
  int* sel = pq + 2 * PM;               // [JN]
  int* live = sel + JN;                 // [JN]
  int* flag = live + JN;                // [8]
  const Layout& L = a.L;
  const int b = blockIdx.y;
  const int slot = a.active[blockIdx.x];
  int* NI = node_int(a, slot);
  if (a.step > 0 && !NI[I_MORE + b]) return;
  double* S = node_ptr(a, slot);
  const int tid = threadIdx.x, lane = tid & 31, p = L.p[b];
  // ---- set-up, stage 1: M (Gram matrix of the residual panel) -> guarded Cholesky-QR factor T.  M at 0, L at RR_MM, T at RR_TM.
  double* Tm = Q + RR_TM;               // [PM][ZLD]
  {
    double* M0 = Q; double* L0 = Q + RR_MM;
    for (int e = tid; e < PM * PM; e += 256) M0[(e / PM) * ZLD + (e % PM)] = S[L.Q[b] + e];
    __syncthreads();
    // residual of the minority columns (before this step; column p - 1 holds the probe) and the scale of H
    double r2 = 0.0, h2 = 0.0;
    if (tid < PM) {
      if (tid < p - 1 && S[L.th[b] + tid] > 0.0) r2 = M0[tid * ZLD + tid];
      double h = 0.0;
      for (int i = 0; i < PM; ++i) { const double v = S[L.H[b] + i * PM + tid]; h += v * v; }
      h2 = h;
    }
    r2 = block_sum(r2, red);
    h2 = block_max(h2, red);
    if (tid == 0) { lam[0] = r2; lam[1] = h2; }
    __syncthreads();
    if (tid < 32) chol_guard_warp(M0, L0, Tm, valid, lane, 1e-10, 1e-20 * lam[1], p, 1e-10, (p > 1) ? p - 1 : -1);
    __syncthreads();
  }
  const double res2 = lam[0];
  // ---- stage 2: X0 = Z'W2, C0 = R'W2 (sums of the row tiles' partials, ld PM) over the dead M; Y = C0 T over the dead L (ld PM)
  double* X0 = Q + RR_X0; double* C0 = Q + RR_C0; double* Ym = Q + RR_MM;
  for (int e = tid; e < PM * PM; e += 256) {
    double sx = 0.0, sc_ = 0.0;
    for (int t = 0; t < L.nt[b]; ++t) {
      sx += S[L.partA[b] + (size_t)t * PM * PM + e];
      sc_ += S[L.partB[b] + (size_t)t * PM * PM + e];
    }
    X0[e] = sx; C0[e] = sc_;
  }
  // A = [[sym(H), X0 T], [., T' C0 T]]
  for (int e = tid; e < JN * JLD; e += 256) A[e] = 0.0;
  __syncthreads();
  for (int e = tid; e < PM * PM; e += 256) {
    const int i = e / PM, j = e % PM;
    A[i * JLD + j] = 0.5 * (S[L.H[b] + i * PM + j] + S[L.H[b] + j * PM + i]);
    double s = 0.0, y = 0.0;
    for (int c = 0; c < PM; ++c) {
      const double t = Tm[c * ZLD + j];
      s = fma(X0[i * PM + c], t, s);
      y = fma(C0[i * PM + c], t, y);
    }
    A[i * JLD + PM + j] = s; A[(PM + j) * JLD + i] = s;
    Ym[i * PM + j] = y;
  }
  __syncthreads();
  for (int e = tid; e < PM * PM; e += 256) {
    const int i = e / PM, j = e % PM;
    double sij = 0.0, sji = 0.0;                  // (T' C0 T)_ij and its transpose entry: the block is symmetrised on the fly
    for (int c = 0; c < PM; ++c) {
      sij = fma(Tm[c * ZLD + i], Ym[c * PM + j], sij);
      sji = fma(Tm[c * ZLD + j], Ym[c * PM + i], sji);
    }
    A[(PM + i) * JLD + PM + j] = 0.5 * (sij + sji);
    S[L.Q[b] + (size_t)PM * PM + e] = Tm[i * ZLD + j];     // park T (the bottom half of Q' is written last)
  }
  __syncthreads();
  double amax = 0.0;
  for (int e = tid; e < JN * JN; e += 256) amax = fmax(amax, fabs(A[(e / JN) * JLD + (e % JN)]));
  amax = block_max(amax, red);
  const double big = 64.0 * (amax + 1.0);
  // ---- compaction: dead columns of Z (index >= p) and dropped residual columns are decoupled from the rest; the Jacobi
  // runs on the live indices only (n2 = their number rounded up to even; a padding index sits at -big).  In steady state
  // most residual columns are converged and dropped, so n2 is close to p instead of 2 PM.
  if (tid == 0) {
    int nl = 0;
    for (int i = 0; i < p; ++i) sel[nl++] = i;
    for (int j = 0; j < PM; ++j) if (valid[j]) sel[nl++] = PM + j;
    flag[1] = nl;
    if (nl & 1) sel[nl++] = -1;
    flag[2] = nl;
  }
  __syncthreads();
  const int nl = flag[1], n2 = flag[2];
  double* Ac = Q;                       // compact matrix (ld JLD)
  for (int e = tid; e < n2 * n2; e += 256) {
    const int i = e / n2, j = e % n2;
    const int li = sel[i], lj = sel[j];
    Ac[i * JLD + j] = (li < 0 || lj < 0) ? ((i == j) ? -big : 0.0) : A[li * JLD + lj];
  }
  __syncthreads();
  if (tid < JN) live[tid] = sel[tid];               // live map (sel is reused for the ranking)
  double* Qc = A;                       // eigenvector accumulator
  for (int e = tid; e < n2 * JLD; e += 256) Qc[e] = 0.0;
  __syncthreads();
  if (tid < n2) Qc[tid * JLD + tid] = 1.0;
  __syncthreads();
  const int npair = n2 >> 1;
  // ---- cyclic Jacobi on Ac (n2 x n2)
  jacobi_cta(Ac, Qc, n2, a.o.jacobi_sweeps, ((NI[I_CONFIRM] || a.it >= a.o.max_iter) ? 1e-14 : 1e-12) * amax, amax, cs, pq, flag, red);
  if (tid < n2) lam[tid] = Ac[tid * JLD + tid];
  __syncthreads();
  // rank by value (descending, ties by index); the padding index sits at the bottom
  if (tid < n2) {
    const double li = lam[tid];
    int rank = 0;
    for (int j = 0; j < n2; ++j) rank += (lam[j] > li || (lam[j] == li && j < tid)) ? 1 : 0;
    sel[rank] = tid;
  }
  __syncthreads();
  // sign convention per kept column: largest-|.| component positive (kept in cs[c])
  if (tid < p) {
    const int src = sel[tid];
    double best = 0.0, sg = 1.0;
    for (int i = 0; i < nl; ++i) {
      const double v = Qc[i * JLD + src];
      if (fabs(v) > best) { best = fabs(v); sg = (v < 0.0) ? -1.0 : 1.0; }
    }
    cs[tid] = sg;
  }
  __syncthreads();
  // Q' (2 PM x PM): column c = eigenvector sel[c] scattered back to the original indices; bottom half multiplied by T.
  // Mm receives the bottom halves first (compact -> original residual index), then T is applied.  The compact matrix is dead:
  // Mm and T (back from where it was parked) take its place.
  double* Mm = Q; double* Tend = Q + PM * ZLD;
  for (int e = tid; e < PM * PM; e += 256) {
    Mm[(e / PM) * ZLD + (e % PM)] = 0.0;
    Tend[(e / PM) * ZLD + (e % PM)] = S[L.Q[b] + (size_t)PM * PM + e];
  }
  for (int e = tid; e < JN * PM; e += 256) if (e < PM * PM) S[L.Q[b] + e] = 0.0;
  __syncthreads();
  for (int e = tid; e < nl * p; e += 256) {
    const int i = e / p, c = e % p;
    const int li = live[i];
    const double v = cs[c] * Qc[i * JLD + sel[c]];
    if (li < PM) S[L.Q[b] + (size_t)li * PM + c] = v;
    else Mm[(li - PM) * ZLD + c] = v;
  }
  __syncthreads();
  for (int e = tid; e < PM * PM; e += 256) {
    const int ii = e / PM, c = e % PM;
    double v = 0.0;
    if (c < p) for (int j = 0; j < PM; ++j) v = fma(Tend[ii * ZLD + j], Mm[j * ZLD + c], v);
    S[L.Q[b] + (size_t)(PM + ii) * PM + c] = v;
  }
  double thn = -1e300;
  if (tid < p) thn = lam[sel[tid]];
  double tn2 = (tid < p) ? thn * thn : 0.0;
  tn2 = block_sum(tn2, red);
  double rposd = (tid < p && thn > 0.0) ? 1.0 : 0.0;
  rposd = block_sum(rposd, red);
  if (tid < PM) S[L.th[b] + tid] = thn;
  if (tid == 0) {
    const double res = sqrt(fmax(res2, 0.0)) / fmax(sqrt(tn2), 1e-300);
    S[L.scal + S_RES + b] = res;
    const int q = (a.step == 0) ? 1 : NI[I_Q + b] + 1;
    NI[I_Q + b] = q;
    NI[I_R + b] = (int)(rposd + 0.5);
    // a node whose termination decision is pending (confirm) tracks at the tight tolerance until the next scheduled check
    // and may use steps_start steps in the iteration of that check (the host launches that many rounds there)
    const bool last = a.it >= a.o.max_iter;
    const bool confirm = NI[I_CONFIRM] || last;
    const bool check_it = (a.it % a.o.check_every == 0) || last;
    const int ns = (a.it == 1) ? a.o.steps_start : 1;
    const double tol = confirm ? a.o.confirm_tol : a.o.track_tol;
    const int qmax = (confirm && check_it) ? a.o.steps_start : a.o.steps_max;
    const bool stop = (q >= ns) && (res <= tol || q >= qmax);
    NI[I_MORE + b] = stop ? 0 : 1;
  }
}

// k_update: Z <- Z Qtop + R Qbot, W <- W Qtop + W2 Qbot (rows of one tile, in place).
constexpr size_t UPD_SMEM = ((size_t)JN * ZLD + 2 * TS * ZLD) * sizeof(double);
__global__ void __launch_bounds__(256) k_update(BigArgs a) {
  extern __shared__ __align__(16) double sm[];
  double* Qs = sm;                   // [JN][ZLD]
  double* Zs = Qs + JN * ZLD;        // [TS][ZLD]
  double* Rs = Zs + TS * ZLD;
  const Layout& L = a.L;
  const int b = blockIdx.z;
  if ((int)blockIdx.x >= L.nt[b]) return;
  const int slot = a.active[blockIdx.y];
  const int* NI = node_int(a, slot);
  // the update runs for every (node, block) that ran this step: k_rr advanced its step counter to step + 1
  if (a.step > 0 && NI[I_Q + b] != a.step + 1) return;
  double* S = node_ptr(a, slot);
  const int N = L.N[b], tid = threadIdx.x, r0 = blockIdx.x * TS;
  for (int e = tid; e < JN * PM; e += 256) Qs[(e / PM) * ZLD + (e % PM)] = S[L.Q[b] + e];
  for (int pass = 0; pass < 2; ++pass) {
    double* A0 = S + (pass ? L.W[b] : L.Z[b]);
    const double* A1 = S + (pass ? L.W2[b] : L.R[b]);
    __syncthreads();
    for (int e = tid; e < TS * PM; e += 256) {
      const int row = e / PM, col = e % PM;
      const bool ok = r0 + row < N;
      Zs[row * ZLD + col] = ok ? A0[(size_t)(r0 + row) * PM + col] : 0.0;
      Rs[row * ZLD + col] = ok ? A1[(size_t)(r0 + row) * PM + col] : 0.0;
    }
    __syncthreads();
    constexpr int CPT = TS * PM / 256;
    const int row = tid / (PM / CPT), cb = (tid % (PM / CPT)) * CPT;
    double acc[CPT];
#pragma unroll
    for (int q = 0; q < CPT; ++q) acc[q] = 0.0;
#pragma unroll 4
    for (int c = 0; c < PM; ++c) {
      const double z = Zs[row * ZLD + c], r = Rs[row * ZLD + c];
#pragma unroll
      for (int q = 0; q < CPT; ++q) acc[q] = fma(z, Qs[c * ZLD + cb + q], fma(r, Qs[(PM + c) * ZLD + cb + q], acc[q]));
    }
    if (r0 + row < N) {
#pragma unroll
      for (int q = 0; q < CPT; ++q) A0[(size_t)(r0 + row) * PM + cb + q] = acc[q];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CholQR of a panel by one CTA (start bases, periodic re-orthonormalisation of Z): P <- P T.
// ------------------------------------------------------------------------------------------------------------------
__device__ void cholqr_cta(double* P, int N, int pcols, double* sm /* >= 3*PM*ZLD doubles */, int* ism /* PM ints */) {
  double* M = sm; double* Lm = M + PM * ZLD; double* Tm = Lm + PM * ZLD;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int e = tid; e < PM * PM; e += nthr) {
    const int ai = e / PM, bi = e % PM;
    double s = 0.0;
    for (int i = 0; i < N; ++i) s = fma(P[(size_t)i * PM + ai], P[(size_t)i * PM + bi], s);
    M[ai * ZLD + bi] = s;
  }
  __syncthreads();
  if (tid < 32) chol_guard_warp(M, Lm, Tm, ism, tid, 1e-14, 0.0, pcols);
  __syncthreads();
  // one row per thread, in place: out[j] = sum_{c <= j} row[c] T[c][j] (T upper triangular) -> go from the last column down
  for (int i = tid; i < N; i += nthr) {
    double* row = P + (size_t)i * PM;
    for (int j = PM - 1; j >= 0; --j) {
      double v = 0.0;
      for (int c = 0; c <= j; ++c) v = fma(row[c], Tm[c * ZLD + j], v);
      row[j] = v;
    }
  }
  __syncthreads();
}

// start bases of the three blocks (same for every node): base[b] = CholQR2(hash panel); block 2 carries the k unit vectors
// of the identity corner first.  grid = 3.
__global__ void __launch_bounds__(256) k_start_basis(Layout L, double* base0, double* base1, double* base2, int seed) {
  __shared__ double sm[3 * PM * ZLD];
  __shared__ int ism[PM];
  const int b = blockIdx.x;
  double* P = (b == 0) ? base0 : ((b == 1) ? base1 : base2);
  const int N = L.N[b], p = L.p[b], n = L.n;
  const int q = (b == 1) ? (L.k < p ? L.k : p) : 0;
  for (int e = threadIdx.x; e < N * PM; e += 256) {
    const int i = e / PM, j = e - i * PM;
    double v = (j < p) ? hash_unit((unsigned long long)i, (unsigned long long)j, (unsigned long long)(seed + b)) : 0.0;
    if (b == 1) {
      if (i >= n && i < n + q) v = 0.0;
      if (j < q) v = (i == n + j) ? 1.0 : 0.0;
    }
    P[e] = v;
  }
  __syncthreads();
  cholqr_cta(P, N, p, sm, ism);
  cholqr_cta(P, N, p, sm, ism);
}

__global__ void __launch_bounds__(256) k_reorth(BigArgs a) {
  __shared__ double sm[3 * PM * ZLD];
  __shared__ int ism[PM];
  const Layout& L = a.L;
  const int b = blockIdx.y;
  const int slot = a.active[blockIdx.x];
  const int* NI = node_int(a, slot);
  double* S = node_ptr(a, slot);
  cholqr_cta(S + L.Z[b], L.N[b], L.p[b], sm, ism);
}

// ------------------------------------------------------------------------------------------------------------------
// node setup: cut table (OMC.jl:1581-1676), gathered cut vectors, Gram matrix of the dense rows, cold
// start (oracle/bigblock.py: BigState).  One CTA per node.
// ------------------------------------------------------------------------------------------------------------------
struct SetupArgs {
  const double* pool_x; const double* pool_vhat;
  const int* cut_ptr; const int* cut_ids; const unsigned char* cut_dirs;
  const double* base[3];
};

__device__ __forceinline__ void cut_coeffs(int cut_type, int dir, double h, int fix3, double& lb, double& ub, double& al, double& be) {
  const double ab = fabs(h);
  if (cut_type == OMC_CUT_LINEAR) {
    if (dir == 0) { lb = -1.0; ub = h; al = h - 1.0; be = h; }
    else { lb = h; ub = 1.0; al = h + 1.0; be = -h; }
  } else if (cut_type == OMC_CUT_LINEAR2) {
    if (dir == 0) { lb = -1.0; ub = -ab; al = -(1.0 + ab); be = -ab; }
    else if (dir == 1) { lb = -ab; ub = ab; al = 0.0; be = h * h; }
    else { lb = ab; ub = 1.0; al = 1.0 + ab; be = -ab; }
  } else {
    if (dir == 0) { lb = -1.0; ub = -ab; al = -(1.0 + ab); be = -ab; }
    else if (dir == 1) { lb = -ab; ub = 0.0; al = -ab; be = 0.0; }
    else if (dir == 2) { lb = 0.0; ub = ab; al = ab; be = 0.0; }
    else if (fix3) { lb = ab; ub = 1.0; al = 1.0 + ab; be = -ab; }
    else { lb = ab; ub = 1.0; al = ab; be = 0.0; }          // OMC.jl:1675 (reference quirk Q1)
  }
}

__global__ void __launch_bounds__(256) k_node_init(BigArgs a, SetupArgs sa_) {
  extern __shared__ __align__(16) double sm[];   // XX [Lc][Lc]
  const Layout& L = a.L;
  const int slot = blockIdx.x;
  double* S = node_ptr(a, slot);
  int* NI = node_int(a, slot);
  const int n = L.n, m = L.m, k = L.k, tid = threadIdx.x;
  const int e0 = sa_.cut_ptr[slot], Lc = sa_.cut_ptr[slot + 1] - e0;
  const int rq = 1 + Lc * (k + 1);
  // (the record was zeroed by a memset on the host side)
  for (int i = tid; i < k; i += 256) S[L.V[1] + (size_t)(n + i) * L.N[1] + n + i] = 1.0;    // E2
  for (int i = tid; i < n; i += 256) S[L.V[2] + (size_t)i * n + i] = a.a;                   // a I
  for (int b = 0; b < 3; ++b) {
    for (int e = tid; e < L.N[b] * PM; e += 256) S[L.Z[b] + e] = sa_.base[b][e];
    for (int c = tid; c < PM; c += 256) {
      double th = (c < L.p[b]) ? 0.0 : -1e300;
      if (b == 1 && c < k && c < L.p[b]) th = 1.0;
      S[L.th[b] + c] = th;
    }
  }
  // cuts
  for (int e = tid; e < Lc * n; e += 256) {
    const int l = e / n, i = e - l * n;
    S[L.xs + e] = sa_.pool_x[(size_t)sa_.cut_ids[e0 + l] * n + i];
  }
  for (int e = tid; e < Lc * k; e += 256) {
    const int l = e / k, j = e - l * k;
    const double h = sa_.pool_vhat[(size_t)sa_.cut_ids[e0 + l] * k + j];
    double lb, ub, al, be;
    cut_coeffs(a.o.cut_type, sa_.cut_dirs[(size_t)(e0 + l) * k + j], h, a.o.fix_linear3_right, lb, ub, al, be);
    S[L.clb + e] = a.sa * lb; S[L.cub + e] = a.sa * ub; S[L.cal + e] = a.sa * al;
    S[L.rhs + e] = a.a * be;     // per-column beta, summed below
  }
  __syncthreads();
  for (int l = tid; l < Lc; l += 256) {
    double be = 0.0;
    for (int j = 0; j < k; ++j) be += S[L.rhs + (size_t)l * k + j];
    S[L.cbe + l] = be;
  }
  __syncthreads();
  // v-form rows of a cold start: U = 0, Y = 0 -> vv = clip(0), vg = max(beta, 0)
  for (int e = tid; e < Lc * k; e += 256) S[L.vv + e] = clampd(0.0, S[L.clb + e], S[L.cub + e]);
  for (int l = tid; l < Lc; l += 256) S[L.vg + l] = fmax(S[L.cbe + l], 0.0);
  // XX = x x'
  for (int e = tid; e < Lc * Lc; e += 256) {
    const int l1 = e / Lc, l2 = e - l1 * Lc;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = fma(S[L.xs + (size_t)l1 * n + i], S[L.xs + (size_t)l2 * n + i], s);
    sm[e] = s;
  }
  __syncthreads();
  // G (rq x rq, ld rcap): rows/cols: 0 trace; 1 + l k + j: v_lj; 1 + L k + l: g_l
  const int ld = L.rcap;
  for (int e = tid; e < rq * rq; e += 256) {
    const int r = e / rq, c = e - r * rq;
    auto kind = [&](int q, int& l, int& j) { if (q == 0) return 0; if (q < 1 + Lc * k) { l = (q - 1) / k; j = (q - 1) - l * k; return 1; } l = q - 1 - Lc * k; j = 0; return 2; };
    int l1 = 0, j1 = 0, l2 = 0, j2 = 0;
    const int k1 = kind(r, l1, j1), k2 = kind(c, l2, j2);
    double v = 0.0;
    if (k1 == 0 && k2 == 0) v = (double)n;
    else if (k1 == 0 && k2 == 2) v = sm[l2 * Lc + l2];
    else if (k1 == 2 && k2 == 0) v = sm[l1 * Lc + l1];
    else if (k1 == 1 && k2 == 1) v = (j1 == j2) ? sm[l1 * Lc + l2] : 0.0;
    else if (k1 == 1 && k2 == 2) v = S[L.cal + (size_t)l2 * k + j1] * sm[l1 * Lc + l2];
    else if (k1 == 2 && k2 == 1) v = S[L.cal + (size_t)l1 * k + j2] * sm[l1 * Lc + l2];
    else if (k1 == 2 && k2 == 2) {
      double s = 0.0;
      for (int j = 0; j < k; ++j) s += S[L.cal + (size_t)l1 * k + j] * S[L.cal + (size_t)l2 * k + j];
      const double xx = sm[l1 * Lc + l2];
      v = s * xx + xx * xx;
    }
    S[L.G + (size_t)r * ld + c] = v;
  }
  if (tid == 0) {
    S[L.scal + S_RHO] = a.o.rho0; S[L.scal + S_V4] = a.ktr; S[L.scal + S_CFAC] = 1.0;
    S[L.scal + S_LB] = -1e300; S[L.scal + S_RP] = 1e300; S[L.scal + S_RD] = 1e300;
    for (int q = 0; q < ISTR; ++q) NI[q] = 0;
    NI[I_STATUS] = -1; NI[I_L] = Lc; NI[I_SLOT] = slot;
  }
  (void)m;
}

// Minv = ((sigma + 3 rho)/rho I + G)^-1 by Gauss-Jordan in global memory (one CTA per node; rebuilt when rho changes).
__global__ void __launch_bounds__(256) k_minv(BigArgs a, int only_adapted) {
  const Layout& L = a.L;
  const int slot = a.active[blockIdx.x];
  double* S = node_ptr(a, slot);
  int* NI = node_int(a, slot);
  if (only_adapted && !NI[I_ADAPTED]) return;
  const int Lc = NI[I_L], k = L.k, rq = 1 + Lc * (k + 1), ld = L.rcap, tid = threadIdx.x;
  const double rho = S[L.scal + S_RHO];
  const double dg = (a.o.sigma + 3.0 * rho) / rho;
  double* M = S + L.Minv;
  for (int e = tid; e < rq * rq; e += 256) {
    const int r = e / rq, c = e - r * rq;
    M[(size_t)r * ld + c] = S[L.G + (size_t)r * ld + c] + ((r == c) ? dg : 0.0);
  }
  __syncthreads();
  for (int p = 0; p < rq; ++p) {
    const double inv = 1.0 / M[(size_t)p * ld + p];
    __syncthreads();
    for (int j = tid; j < rq; j += 256) if (j != p) M[(size_t)p * ld + j] *= inv;
    __syncthreads();
    for (int e = tid; e < rq * rq; e += 256) {
      const int i = e / rq, j = e - i * rq;
      if (i != p && j != p) M[(size_t)i * ld + j] -= M[(size_t)i * ld + p] * M[(size_t)p * ld + j];
    }
    __syncthreads();
    for (int i = tid; i < rq; i += 256) if (i != p) M[(size_t)i * ld + p] *= -inv;
    if (tid == 0) M[(size_t)p * ld + p] = inv;
    __syncthreads();
  }
  if (tid == 0) NI[I_ADAPTED] = 0;
}

// ------------------------------------------------------------------------------------------------------------------
// k_check: residuals and objective pieces, per tile.  chk_part[tile][q]:
//   0 rp (max), 1 rd (max), 2 n_p (max), 3 n_d (max), 4 sum Mk X^2, 5 sum Mk (X - A)^2, 6 tr T, 7 tr mu2 corner / rho,
//   8 tr(F3) (-> a tr mu3 = -rho a tr F3), 9 box dual sum / rho.   rows_part gets the dense rows of the CURRENT (Y, U).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_check(BigArgs a) {
  extern __shared__ __align__(16) double sm[];
  const Layout& L = a.L;
  const int slot = a.active[blockIdx.y];
  double* S = node_ptr(a, slot);
  const int* NI = node_int(a, slot);
  const int n = L.n, m = L.m, k = L.k, N1 = L.N[0], N2 = L.N[1];
  const int Lc = NI[I_L];
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* Zi_[3] = {sm + YOFF_I0, sm + YOFF_I1, sm + YOFF_I2};
  double* Zj_[3] = {sm + YOFF_J0, sm + YOFF_J1, sm + YOFF_J2};
  const int ld_[3] = {ZLD, ZLD, YLD2};
  int nc_[3];
  double* sc = sm + YOFF_SC;
  double* ub = sm + YOFF_UB;
  double* xi = sm + YOFF_X;
  double* xj = xi + (size_t)L.Lcap * TS;
  double* mgs = xj + (size_t)L.Lcap * TS;   // [Lc] mg / rho
  double* wred = mgs + L.Lcap + 1;          // [8][rcap]
  double* red = wred + 8 * (size_t)L.rcap;  // [32]
  const double rho = S[L.scal + S_RHO];
  const int t = blockIdx.x;
  const int nX = L.tn * L.tm, nT = L.tm * (L.tm + 1) / 2, nYt = L.tn * (L.tn + 1) / 2;
  for (int e = threadIdx.x; e < 8 * L.rcap; e += 256) wred[e] = 0.0;
  for (int b = 0; b < 3; ++b) nc_[b] = stage_scale(sc + b * PM, S + L.th[b], L.p[b]);
  for (int l = threadIdx.x; l < Lc; l += 256) mgs[l] = S[L.vg + l] - fmax(S[L.vg + l], 0.0);
  __syncthreads();
  double rp = 0.0, rd = 0.0, np_ = 0.0, nd_ = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0, s8 = 0.0, s9 = 0.0;
  double F[4][4];
  double* part = S + L.chk_part + (size_t)t * NCHK;
  if (t < nX) {
    const int I = t / L.tm, J = t - I * L.tm, i0 = I * TS, j0 = J * TS;
    stage_panel(Zi_[0], ZLD, S + L.Z[0], i0, n, sc, nc_[0]);
    stage_panel(Zj_[0], ZLD, S + L.Z[0] + (size_t)n * PM, j0, m, sc, nc_[0]);
    __syncthreads();
    lowrank_tile(F, Zi_[0], Zj_[0], ZLD, nc_[0], ty, tx);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = i0 + ty + 16 * r, j = j0 + tx + 16 * c;
        if (i < n && j < m) {
          const size_t e = (size_t)i * m + j;
          const double x = S[L.X + e], mk = (double)a.Mk[e], am = a.AM[e];
          const double mu = rho * (S[L.V[0] + (size_t)i * N1 + n + j] - F[r][c]);
          rp = fmax(rp, fabs(x - F[r][c]));
          np_ = fmax(np_, fabs(F[r][c]));
          const double gX = -2.0 * mu;
          if (a.sh.on) {                    // Shor form: linear objective; the X_t stationarity is finished by k_shor_check_c
            a.sh.SS[(size_t)slot * a.sh.SL.total + a.sh.SL.Xtt + e] = gX;
            nd_ = fmax(nd_, fmax(fabs(am), fabs(gX)));
            s5 += am * x;                   // sum_Omega A X
          } else {
            rd = fmax(rd, fabs(mk * x - am - gX));
            nd_ = fmax(nd_, fmax(fabs(mk * x), fmax(fabs(am), fabs(gX))));
            s4 += mk * x * x;
            const double d = mk * x - am;     // mk (x - A)
            s5 += d * d;
          }
        }
      }
  } else if (t < nX + nT) {
    int I, J;
    lower_tile(t - nX, I, J);
    const int i0 = I * TS, j0 = J * TS;
    stage_panel(Zi_[0], ZLD, S + L.Z[0] + (size_t)n * PM, i0, m, sc, nc_[0]);
    stage_panel(Zj_[0], ZLD, S + L.Z[0] + (size_t)n * PM, j0, m, sc, nc_[0]);
    __syncthreads();
    lowrank_tile(F, Zi_[0], Zj_[0], ZLD, nc_[0], ty, tx);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = i0 + ty + 16 * r, j = j0 + tx + 16 * c;
        if (i < m && j < m) {
          const double tt = S[L.T + (size_t)i * m + j];
          const double mu = rho * (S[L.V[0] + (size_t)(n + i) * N1 + n + j] - F[r][c]);
          rp = fmax(rp, fabs(tt - F[r][c]));
          np_ = fmax(np_, fabs(F[r][c]));
          const double gT = -mu;
          if (a.sh.on && i == j) a.sh.SS[(size_t)slot * a.sh.SL.total + a.sh.SL.gTd + j] = gT;     // finished by k_shor_check_col
          else rd = fmax(rd, fabs(((i == j) ? a.cT : 0.0) - gT));
          nd_ = fmax(nd_, fabs(gT));
          if (i == j) s6 += tt;
        }
      }
  } else if (t < nX + nT + nYt) {
    int I, J;
    lower_tile(t - nX - nT, I, J);
    const int i0 = I * TS, j0 = J * TS;
    for (int b = 0; b < 3; ++b) {
      stage_panel(Zi_[b], ld_[b], S + L.Z[b], i0, n, sc + b * PM, nc_[b]);
      stage_panel(Zj_[b], ld_[b], S + L.Z[b], j0, n, sc + b * PM, nc_[b]);
    }
    for (int e = threadIdx.x; e < Lc * TS; e += 256) {
      const int l = e / TS, q = e - l * TS;
      xi[l * TS + q] = (i0 + q < n) ? S[L.xs + (size_t)l * n + i0 + q] : 0.0;
      xj[l * TS + q] = (j0 + q < n) ? S[L.xs + (size_t)l * n + j0 + q] : 0.0;
    }
    __syncthreads();
    double F2[4][4], F3[4][4], yv[4][4];
    lowrank_tile(F, Zi_[0], Zj_[0], ld_[0], nc_[0], ty, tx);
    lowrank_tile(F2, Zi_[1], Zj_[1], ld_[1], nc_[1], ty, tx);
    lowrank_tile(F3, Zi_[2], Zj_[2], ld_[2], nc_[2], ty, tx);
    const double v4 = S[L.scal + S_V4];
    const double m4 = rho * (v4 - fmax(v4, 0.0));
    double trp = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int il = ty + 16 * r, jl = tx + 16 * c, i = i0 + il, j = j0 + jl;
        yv[r][c] = 0.0;
        if (i < n && j < n) {
          const double y = S[L.Y + (size_t)i * n + j];
          yv[r][c] = y;
          const double V3 = S[L.V[2] + (size_t)i * n + j];
          const double s3 = V3 + F3[r][c];
          const double z3 = ((i == j) ? a.a : 0.0) - y;
          rp = fmax(rp, fmax(fabs(y - F[r][c]), fmax(fabs(y - F2[r][c]), fabs(z3 - s3))));
          np_ = fmax(np_, fmax(fabs(F[r][c]), fmax(fabs(F2[r][c]), fabs(s3))));
          const double mu1 = rho * (S[L.V[0] + (size_t)i * N1 + j] - F[r][c]);
          const double mu2 = rho * (S[L.V[1] + (size_t)i * N2 + j] - F2[r][c]);
          const double mu3 = -rho * F3[r][c];
          double g = -mu1 - mu2 + mu3 + ((i == j) ? m4 : 0.0);
          for (int l = 0; l < Lc; ++l) g = fma(rho * mgs[l], xi[l * TS + il] * xj[l * TS + jl], g);
          rd = fmax(rd, fabs(g));
          nd_ = fmax(nd_, fabs(g));
          if (i == j) { s8 += F3[r][c]; trp += y; }
        }
      }
    const double wgt = (I != J) ? 2.0 : 1.0;
    trp = warp_sum(trp);
    if (lane == 0) wred[warp * L.rcap + 0] = trp;
    for (int l = 0; l < Lc; ++l) {
      double q = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) q = fma(xi[l * TS + ty + 16 * r] * xj[l * TS + tx + 16 * c], yv[r][c], q);
      q = warp_sum(q);
      if (lane == 0) wred[warp * L.rcap + 1 + Lc * k + l] = wgt * q;
    }
    __syncthreads();
    double* rpart = S + L.rows_part + (size_t)(t - nX - nT) * L.rcap;
    for (int q = threadIdx.x; q < 1 + Lc * (k + 1); q += 256) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += wred[w * L.rcap + q];
      rpart[q] = s;
    }
  } else {
    const int ut = t - nX - nT - nYt;
    const int i0 = ut * TS;
    double* Zi = Zi_[1]; double* Zj = Zj_[1];
    const int nc1 = nc_[1];
    stage_panel(Zi, ZLD, S + L.Z[1], i0, n, sc + PM, nc1);
    for (int e = threadIdx.x; e < k * PM; e += 256) {
      const int j = e / PM, c = e - j * PM;
      Zj[j * ZLD + c] = S[L.Z[1] + (size_t)(n + j) * PM + c] * sc[PM + c];
    }
    for (int e = threadIdx.x; e < Lc * TS; e += 256) {
      const int l = e / TS, q = e - l * TS;
      xi[l * TS + q] = (i0 + q < n) ? S[L.xs + (size_t)l * n + i0 + q] : 0.0;
    }
    __syncthreads();
    const double sa = a.sa;
    for (int e = threadIdx.x; e < TS * k; e += 256) {
      const int il = e / k, j = e - il * k, i = i0 + il;
      double u = 0.0;
      if (i < n) {
        double Fv = 0.0;
        for (int c = 0; c < nc1; ++c) Fv = fma(Zi[il * ZLD + c], Zj[j * ZLD + c], Fv);
        u = S[L.U + (size_t)i * k + j];
        const double v5 = S[L.v5 + (size_t)i * k + j];
        const double lo5 = (i >= n - k + j) ? 0.0 : -sa;
        const double s5v = clampd(v5, lo5, sa);
        const double m5 = rho * (v5 - s5v);
        const double mu2 = rho * (S[L.V[1] + (size_t)i * N2 + n + j] - Fv);
        rp = fmax(rp, fmax(fabs(u - Fv), fabs(u - s5v)));
        np_ = fmax(np_, fmax(fabs(Fv), fabs(s5v)));
        double g = -2.0 * mu2 - m5;
        for (int l = 0; l < Lc; ++l) {
          const double vv = S[L.vv + (size_t)l * k + j];
          const double mv = rho * (vv - clampd(vv, S[L.clb + (size_t)l * k + j], S[L.cub + (size_t)l * k + j]));
          g -= xi[l * TS + il] * (mv + rho * mgs[l] * S[L.cal + (size_t)l * k + j]);
        }
        rd = fmax(rd, fabs(g));
        nd_ = fmax(nd_, fabs(g));
        s9 += (m5 < 0.0) ? m5 * lo5 : m5 * sa;
      }
      ub[il * 17 + j] = u;
    }
    __syncthreads();
    double* rpart = S + L.rows_part + (size_t)(nYt + ut) * L.rcap;
    for (int e = threadIdx.x; e < Lc * k; e += 256) {
      const int l = e / k, j = e - l * k;
      double q = 0.0;
      for (int il = 0; il < TS; ++il) q = fma(xi[l * TS + il], ub[il * 17 + j], q);
      rpart[1 + l * k + j] = q;
    }
    if (threadIdx.x == 0) rpart[0] = 0.0;
    for (int l = threadIdx.x; l < Lc; l += 256) rpart[1 + Lc * k + l] = 0.0;
    if (ut == 0) {
      // identity corner of block 2: z = delta, s = F, mu = rho (V - F)
      for (int e = threadIdx.x; e < k * k; e += 256) {
        const int i = e / k, j = e - i * k;
        double Fv = 0.0;
        for (int c = 0; c < nc1; ++c) Fv = fma(Zj[i * ZLD + c], Zj[j * ZLD + c], Fv);
        rp = fmax(rp, fabs(((i == j) ? 1.0 : 0.0) - Fv));
        np_ = fmax(np_, fabs(Fv));
        if (i == j) s7 += S[L.V[1] + (size_t)(n + i) * N2 + n + j] - Fv;
      }
    }
  }
  rp = block_max(rp, red); rd = block_max(rd, red); np_ = block_max(np_, red); nd_ = block_max(nd_, red);
  s4 = block_sum(s4, red); s5 = block_sum(s5, red); s6 = block_sum(s6, red); s7 = block_sum(s7, red);
  s8 = block_sum(s8, red); s9 = block_sum(s9, red);
  if (threadIdx.x == 0) {
    part[0] = rp; part[1] = rd; part[2] = np_; part[3] = nd_; part[4] = s4; part[5] = s5; part[6] = s6; part[7] = s7;
    part[8] = s8; part[9] = s9;
  }
}

// k_decide: one warp-CTA per node: reduce, decide, adapt rho.
__global__ void __launch_bounds__(128) k_decide(BigArgs a, int* counters /* [0] active after, [1] confirm, [2] adapted */) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double red[32];
  const Layout& L = a.L;
  const int slot = a.active[blockIdx.x];
  double* S = node_ptr(a, slot);
  int* NI = node_int(a, slot);
  const int k = L.k, n = L.n, m = L.m, Lc = NI[I_L], tid = threadIdx.x;
  const int was_confirm = NI[I_CONFIRM];
  const int rq = 1 + Lc * (k + 1);
  double* R0 = sm;
  for (int q = tid; q < rq; q += 128) {
    double s = 0.0;
    for (int t = 0; t < L.tilesYU; ++t) s += S[L.rows_part + (size_t)t * L.rcap + q];
    R0[q] = s;
  }
  double rp = 0.0, rd = 0.0, np_ = 0.0, nd_ = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0, s8 = 0.0, s9 = 0.0;
  for (int t = tid; t < L.tilesAll; t += 128) {
    const double* p = S + L.chk_part + (size_t)t * NCHK;
    rp = fmax(rp, p[0]); rd = fmax(rd, p[1]); np_ = fmax(np_, p[2]); nd_ = fmax(nd_, p[3]);
  }
  // sums in fixed order by one thread each (deterministic)
  if (tid < 6) {
    double s = 0.0;
    for (int t = 0; t < L.tilesAll; ++t) s += S[L.chk_part + (size_t)t * NCHK + 4 + tid];
    red[tid] = s;
  }
  __syncthreads();
  s4 = red[0]; s5 = red[1]; s6 = red[2]; s7 = red[3]; s8 = red[4]; s9 = red[5];
  __syncthreads();
  const double rho = S[L.scal + S_RHO];
  // scalar rows
  double dsum = 0.0, bmax = 0.0;
  for (int e = tid; e < Lc * k; e += 128) {
    const double vv = S[L.vv + e], lb = S[L.clb + e], ub = S[L.cub + e];
    const double sv = clampd(vv, lb, ub), mv = rho * (vv - sv);
    rp = fmax(rp, fabs(R0[1 + e] - sv));
    np_ = fmax(np_, fabs(sv));
    dsum -= (mv < 0.0) ? mv * lb : mv * ub;
  }
  for (int l = tid; l < Lc; l += 128) {
    const double vg = S[L.vg + l], sg = fmax(vg, 0.0), mg = rho * (vg - sg), be = S[L.cbe + l];
    double zg = be - R0[1 + Lc * k + l];
    for (int j = 0; j < k; ++j) zg += S[L.cal + (size_t)l * k + j] * R0[1 + l * k + j];
    rp = fmax(rp, fabs(zg - sg));
    np_ = fmax(np_, fmax(fabs(sg), fabs(be)));
    bmax = fmax(bmax, fabs(be));
    dsum += be * mg;
  }
  rp = block_max(rp, red); rd = block_max(rd, red); np_ = block_max(np_, red); nd_ = block_max(nd_, red);
  dsum = block_sum(dsum, red);
  if (tid != 0) return;
  const double v4 = S[L.scal + S_V4], s4r = fmax(v4, 0.0), m4 = rho * (v4 - s4r);
  rp = fmax(rp, fabs((a.ktr - R0[0]) - s4r));
  np_ = fmax(np_, fmax(fabs(s4r), fmax(a.a, a.ktr)));
  nd_ = fmax(nd_, a.cT);
  double dual = -0.5 * s4 + a.c0 + rho * s7 - rho * a.a * s8 + a.ktr * m4 + dsum - s9;
  double objp = 0.5 * s5 + a.cT * s6;
  if (a.sh.on) {      // Shor form (OMC.jl:1838-1846, 1961-1968): 1/2 sum_I (A^2 - 2 A X + W) + cT tr Theta~; rows' residuals from k_shor_check_*
    const double* Qc = a.sh.SS + (size_t)slot * a.sh.SL.total + a.sh.SL.chk;
    rd = fmax(rd, Qc[0]); rp = fmax(rp, Qc[1]);
    dual += Qc[2];
    objp = a.c0 - s5 + 0.5 * Qc[3] + a.cT * s6;
  }
  const double ub = (a.o.cutoff < 1e299) ? a.o.cutoff : 2.0 * fmax(fabs(objp), fabs(dual)) + 1.0;
  auto w1 = [&](double trTb) { return (double)n * a.ktr + sqrt((double)n * m * a.ktr * trTb) + (double)m * trTb + (double)n * k * a.sa; };
  const double bound = dual - rd * w1(ub / a.cT);
  // infeasibility: a feasible node holds the point X = 0, Theta~ = 0 (and W = 0, every Shor variable 0) of value c0.  Without Shor
  // rows the optimum w* has tr Theta~ <= c0 / cT; with them the test point itself (||w||_1 <= w1(0)) bounds the dual value.
  const double bound_c0 = dual - rd * w1(a.sh.on ? 0.0 : a.c0 / a.cT);
  S[L.scal + S_RP] = rp; S[L.scal + S_RD] = rd; S[L.scal + S_OBJP] = objp; S[L.scal + S_OBJD] = dual;
  S[L.scal + S_NP] = np_; S[L.scal + S_ND] = nd_;
  NI[I_ITERS] = a.it;
  NI[I_CONFIRM] = 0;
  bool guard_ok = true;
  double resmax = 0.0;
  for (int b = 0; b < 3; ++b) { guard_ok = guard_ok && (NI[I_R + b] < L.p[b] || L.p[b] == L.N[b]); resmax = fmax(resmax, S[L.scal + S_RES + b]); }
  // mu lies in the dual cone only when the trackers hold every eigenvalue of the minority side (a guard column is left) and
  // are converged: the certified bound is reported from such checks only (it never decreases)
  const bool tracked_ok = (was_confirm || a.it >= a.o.max_iter) && guard_ok && resmax <= 10.0 * a.o.confirm_tol;
  if (tracked_ok && !a.sh.on) S[L.scal + S_LB] = fmax(S[L.scal + S_LB], bound);
  int decision = -1;
  if (!(fabs(objp) < 1e300 && fabs(dual) < 1e300 && rp < 1e300 && rd < 1e300)) {
    NI[I_STATUS] = OMC_STATUS_NUMERICAL; NI[I_DONE] = 1;
    return;
  }
  // (with Shor rows ||w*||_1 is not bounded by the formula above: no certified bound and no cut-off)
  if (rp <= a.o.eps_abs + a.o.eps_rel * np_ && rd <= a.o.eps_abs + a.o.eps_rel * nd_) decision = OMC_STATUS_OPTIMAL;
  else if (!a.sh.on && a.o.cutoff < 1e299 && bound > a.o.cutoff) decision = OMC_STATUS_CUTOFF;
  else if (a.o.infeasible_by_bound && Lc > 0 && bound_c0 > a.c0 * (1.0 + 1e-9) + 1e-12) decision = OMC_STATUS_INFEASIBLE;
  if (decision >= 0) {
    if (tracked_ok || a.it >= a.o.max_iter) {
      NI[I_STATUS] = tracked_ok ? decision : OMC_STATUS_ITERATION_LIMIT;
      NI[I_DONE] = 1;
      return;
    }
    NI[I_CONFIRM] = 1;
    atomicAdd(&counters[0], 1);
    atomicAdd(&counters[1], 1);
    return;
  }
  if (a.it >= a.o.max_iter) { NI[I_STATUS] = OMC_STATUS_ITERATION_LIMIT; NI[I_DONE] = 1; return; }
  atomicAdd(&counters[0], 1);
  if (a.o.adapt_every > 0 && (a.it % a.o.adapt_every == 0)) {
    const double ratio = sqrt((rp / fmax(np_, 1e-12)) / fmax(rd / fmax(nd_, 1e-12), 1e-30));
    if (ratio > a.o.adapt_thresh || ratio < 1.0 / a.o.adapt_thresh) {
      const double rho_new = fmin(fmax(rho * ratio, 1e-6), 1e6);
      S[L.scal + S_CFAC] = rho / rho_new;
      S[L.scal + S_RHO] = rho_new;
      NI[I_ADAPTED] = 1;
      atomicAdd(&counters[2], 1);
    }
  }
}

// k_rescale: after a rho change, mu stays: v <- s + cfac (v - s) on every cone row; eigenvectors are unchanged and the
// Ritz values of the mu part scale by cfac.  grid (tiles of 64 rows of V_b, node, b); b = 3: scalar rows + theta.
__global__ void __launch_bounds__(256) k_rescale(BigArgs a) {
  __shared__ double sc[PM];
  __shared__ double Zr[PM];
  const Layout& L = a.L;
  const int slot = a.active[blockIdx.y];
  const int* NI = node_int(a, slot);
  if (!NI[I_ADAPTED]) return;
  double* S = node_ptr(a, slot);
  const double cf = S[L.scal + S_CFAC];
  const int b = blockIdx.z, tid = threadIdx.x;
  if (b == 3) {
    if (blockIdx.x != 0) return;
    const int k = L.k, n = L.n, Lc = NI[I_L];
    const double sa = a.sa;
    if (tid == 0) { const double v4 = S[L.scal + S_V4], s4 = fmax(v4, 0.0); S[L.scal + S_V4] = s4 + cf * (v4 - s4); }
    for (int e = tid; e < n * k; e += 256) {
      const int i = e / k, j = e - i * k;
      const double v = S[L.v5 + e], lo5 = (i >= n - k + j) ? 0.0 : -sa, s = clampd(v, lo5, sa);
      S[L.v5 + e] = s + cf * (v - s);
    }
    for (int e = tid; e < Lc * k; e += 256) {
      const double v = S[L.vv + e], s = clampd(v, S[L.clb + e], S[L.cub + e]);
      S[L.vv + e] = s + cf * (v - s);
    }
    for (int l = tid; l < Lc; l += 256) { const double v = S[L.vg + l], s = fmax(v, 0.0); S[L.vg + l] = s + cf * (v - s); }
    return;
  }
  if ((int)blockIdx.x >= L.nt[b]) return;
  const int N = L.N[b];
  if (tid < PM) { const double th = S[L.th[b] + tid]; sc[tid] = (tid < L.p[b] && th > 0.0) ? th : 0.0; }
  __syncthreads();
  // side +: V <- F + cf (V - F) = cf V + (1 - cf) F;   side - (b = 2): s = V + Fm -> V <- V + (1 - cf) Fm ... with mu/rho = -Fm:
  //         V = s - Fm -> s - cf Fm = V + (1 - cf) Fm
  const double cV = (b == 2) ? 1.0 : cf, cF = 1.0 - cf;
  const int r0 = blockIdx.x * TS;
  for (int row = r0; row < r0 + TS && row < N; ++row) {
    __syncthreads();
    if (tid < PM) Zr[tid] = S[L.Z[b] + (size_t)row * PM + tid] * sc[tid];
    __syncthreads();
    for (int c = tid; c < N; c += 256) {
      double F = 0.0;
#pragma unroll
      for (int q = 0; q < PM; ++q) F = fma(Zr[q], S[L.Z[b] + (size_t)c * PM + q], F);
      const size_t e = L.V[b] + (size_t)row * N + c;
      S[e] = cV * S[e] + cF * F;
    }
  }
}
// theta after a rho change (separate tiny kernel so that k_rescale reads the old values everywhere)
__global__ void __launch_bounds__(3 * PM) k_rescale_theta(BigArgs a) {
  const Layout& L = a.L;
  const int slot = a.active[blockIdx.x];
  const int* NI = node_int(a, slot);
  if (!NI[I_ADAPTED]) return;
  double* S = node_ptr(a, slot);
  const double cf = S[L.scal + S_CFAC];
  const int b = threadIdx.x / PM, c = threadIdx.x % PM;
  if (b >= 3 || c >= L.p[b]) return;
  const double th = S[L.th[b] + c];
  // side +: positive Ritz values are s (unchanged), the others are mu / rho (scaled); side -: the other way round
  if (b < 2) S[L.th[b] + c] = (th > 0.0) ? th : cf * th;
  else S[L.th[b] + c] = (th > 0.0) ? cf * th : th;
}

// k_extract: results in the reference's layout (column-major X n x m, Y n x n, U n x k, unscaled).
__global__ void __launch_bounds__(256) k_extract(BigArgs a, int B, double* outX, double* outY, double* outU, double* outT, int* status, int* iters,
                                                 double* objective, double* lower_bound, double* res) {
  const Layout& L = a.L;
  const int slot = blockIdx.x;
  if (slot >= B) return;
  const double* S = node_ptr(a, slot);
  const int* NI = node_int(a, slot);
  const int n = L.n, m = L.m, k = L.k, tid = threadIdx.x;
  if (outX) for (size_t e = tid; e < (size_t)n * m; e += 256) { const int j = e / n, i = e - (size_t)j * n; outX[(size_t)slot * n * m + e] = S[L.X + (size_t)i * m + j]; }
  if (outY) for (size_t e = tid; e < (size_t)n * n; e += 256) outY[(size_t)slot * n * n + e] = S[L.Y + e] / a.a;
  if (outU) for (size_t e = tid; e < (size_t)n * k; e += 256) { const int j = e / n, i = e - (size_t)j * n; outU[(size_t)slot * n * k + e] = S[L.U + (size_t)i * k + j] / a.sa; }
  if (outT) for (size_t e = tid; e < (size_t)m * m; e += 256) outT[(size_t)slot * m * m + e] = S[L.T + e] * a.a;
  if (tid == 0) {
    const int st = NI[I_STATUS];
    status[slot] = st;
    iters[slot] = NI[I_ITERS];
    lower_bound[slot] = S[L.scal + S_LB];
    objective[slot] = (st == OMC_STATUS_CUTOFF) ? S[L.scal + S_LB] : S[L.scal + S_OBJP];
    res[2 * slot] = S[L.scal + S_RP]; res[2 * slot + 1] = S[L.scal + S_RD];
  }
}


// ------------------------------------------------------------------------------------------------------------------
// k_lanczos: separation oracle for n beyond the in-SM eigensolver (omc_eigsep.cuh, n <= 104): the smallest one or two
// eigenpairs of M = U U' - Y (OMC.jl:2466-2477; feasibility test lambda_min >= -1e-6, OMC.jl:1272-1277) by restarted
// Lanczos with full reorthogonalisation on Aop = Y - U U' (largest of Aop = - smallest of M).  One CTA per node; the
// matrix-vector product never forms U U' (y = Y v - U (U'v)): one warp per row of Y, lanes along the row, warp-shuffle
// reduction.  LM = 48 steps per cycle; the LM x LM tridiagonal is diagonalised by the CTA Jacobi; a cycle restarts from the
// Ritz vector until |beta_m s_m| <= tol |theta|.  The second pair (nev = 2) is found by a second run deflated against the
// first vector.  Y, U column-major per node (Y symmetric).  ws: per node (LM + 3) * n doubles.
// ------------------------------------------------------------------------------------------------------------------
constexpr int LM = 48;
constexpr size_t LANCZOS_SMEM_FIXED = ((size_t)2 * LM * JLD + 4 * LM + 2 * PM + 64 + 16) * sizeof(double) + (2 * PM + 16) * sizeof(int);

__global__ void __launch_bounds__(256) k_lanczos(int n, int k, int B, const double* __restrict__ Y, const double* __restrict__ U, int nev,
                                                 double* __restrict__ ws, double* __restrict__ lam_out, double* __restrict__ vec_out,
                                                 double* __restrict__ bp_out, int* __restrict__ feas_out, double tol, int max_cycles) {
  extern __shared__ __align__(16) double sm[];
  double* T = sm;                        // [LM][JLD]
  double* E = T + LM * JLD;              // [LM][JLD] eigenvectors of T
  double* alpha = E + LM * JLD;          // [LM]
  double* beta = alpha + LM;             // [LM + 1]
  double* coef = beta + LM + 1;          // [LM + 2]
  double* cs = coef + LM + 2;            // [2 PM]
  double* red = cs + 2 * PM;             // [64]
  double* tk = red + 64;                 // [16] U'v
  double* vs = tk + 16;                  // [n]  current vector
  int* pq = reinterpret_cast<int*>(vs + n);
  int* flag = pq + 2 * PM;
  const int node = blockIdx.x;
  if (node >= B) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Yn = Y + (size_t)node * n * n;
  const double* Un = U + (size_t)node * n * k;
  double* Qv = ws + (size_t)node * (LM + 3) * n;     // Lanczos vectors q_0 .. q_LM
  double* X0 = Qv + (size_t)(LM + 1) * n;             // converged vectors
  double* X1 = X0 + n;
  double theta[2] = {0.0, 0.0};
  for (int want = 0; want < nev; ++want) {
    double* Xw = want ? X1 : X0;
    // start vector: deterministic pseudo-random, deflated
    for (int i = tid; i < n; i += 256) vs[i] = hash_unit((unsigned long long)i, (unsigned long long)(7 + want), 777ull);
    __syncthreads();
    double th = 0.0, resid = 1e300;
    for (int cyc = 0; cyc < max_cycles; ++cyc) {
      // q_0 = normalised (deflated) vs
      if (want == 1) {
        double d = 0.0;
        for (int i = tid; i < n; i += 256) d += X0[i] * vs[i];
        d = block_sum(d, red);
        for (int i = tid; i < n; i += 256) vs[i] -= d * X0[i];
        __syncthreads();
      }
      double nr = 0.0;
      for (int i = tid; i < n; i += 256) nr += vs[i] * vs[i];
      nr = sqrt(block_sum(nr, red));
      for (int i = tid; i < n; i += 256) { vs[i] /= nr; Qv[i] = vs[i]; }
      __syncthreads();
      int mdone = 0;
      for (int j = 0; j < LM; ++j) {
        // w = Aop q_j  -> Qv[j + 1]
        if (warp < 8) {
          for (int c = warp; c < k; c += 8) {
            double d = 0.0;
            for (int i = lane; i < n; i += 32) d += Un[(size_t)c * n + i] * vs[i];
            d = warp_sum(d);
            if (lane == 0) tk[c] = d;
          }
        }
        __syncthreads();
        double* w = Qv + (size_t)(j + 1) * n;
        for (int i = warp; i < n; i += 8) {
          const double* yr = Yn + (size_t)i * n;
          double d = 0.0;
          for (int c = lane; c < n; c += 32) d = fma(yr[c], vs[c], d);
          d = warp_sum(d);
          if (lane == 0) {
            for (int c = 0; c < k; ++c) d -= Un[(size_t)c * n + i] * tk[c];
            w[i] = d;
          }
        }
        __syncthreads();
        // alpha_j, full reorthogonalisation against q_0..q_j (and the deflated vector), twice
        for (int pass = 0; pass < 2; ++pass) {
          for (int q = warp; q <= j + want; q += 8) {
            const double* qq = (q <= j) ? Qv + (size_t)q * n : X0;
            double d = 0.0;
            for (int i = lane; i < n; i += 32) d = fma(qq[i], w[i], d);
            d = warp_sum(d);
            if (lane == 0) coef[q] = d;
          }
          __syncthreads();
          if (pass == 0 && tid == 0) alpha[j] = coef[j];
          for (int i = tid; i < n; i += 256) {
            double v = w[i];
            for (int q = 0; q <= j; ++q) v = fma(-coef[q], Qv[(size_t)q * n + i], v);
            if (want == 1) v = fma(-coef[j + 1], X0[i], v);
            w[i] = v;
          }
          __syncthreads();
          if (pass == 1 && tid == 0) alpha[j] += coef[j];
        }
        double bn = 0.0;
        for (int i = tid; i < n; i += 256) bn += w[i] * w[i];
        bn = sqrt(block_sum(bn, red));
        if (tid == 0) beta[j + 1] = bn;
        mdone = j + 1;
        if (bn < 1e-13) break;              // invariant subspace found
        for (int i = tid; i < n; i += 256) { const double v = w[i] / bn; w[i] = v; vs[i] = v; }
        __syncthreads();
      }
      // T (mdone x mdone, padded to even) -> Jacobi
      const int mm = mdone, n2 = mm + (mm & 1);
      for (int e = tid; e < n2 * JLD; e += 256) { T[e] = 0.0; E[e] = 0.0; }
      __syncthreads();
      if (tid < n2) {
        E[tid * JLD + tid] = 1.0;
        if (tid < mm) {
          T[tid * JLD + tid] = alpha[tid];
          if (tid + 1 < mm) { T[tid * JLD + tid + 1] = beta[tid + 1]; T[(tid + 1) * JLD + tid] = beta[tid + 1]; }
        } else T[tid * JLD + tid] = -1e300;     // padding index at the bottom
      }
      __syncthreads();
      double amax = 0.0;
      if (tid < mm) amax = fmax(fabs(alpha[tid]), (tid + 1 < mm) ? fabs(beta[tid + 1]) : 0.0);
      amax = block_max(amax, red);
      if (n2 > mm && tid == 0) T[(n2 - 1) * JLD + n2 - 1] = -64.0 * (amax + 1.0);
      __syncthreads();
      jacobi_cta(T, E, n2, 30, 1e-15 * amax, amax, cs, pq, flag, red);
      // largest Ritz value (the padding sits far below)
      if (tid == 0) {
        int best = 0;
        for (int i = 1; i < mm; ++i) if (T[i * JLD + i] > T[best * JLD + best]) best = i;
        flag[1] = best;
      }
      __syncthreads();
      const int best = flag[1];
      th = T[best * JLD + best];
      resid = fabs(beta[mm] * E[(mm - 1) * JLD + best]);
      // Ritz vector -> vs (and Xw)
      for (int i = tid; i < n; i += 256) {
        double v = 0.0;
        for (int q = 0; q < mm; ++q) v = fma(E[q * JLD + best], Qv[(size_t)q * n + i], v);
        vs[i] = v;
      }
      __syncthreads();
      if (resid <= tol * fmax(1.0, fabs(th)) || mm < LM) break;
    }
    double nr = 0.0;
    for (int i = tid; i < n; i += 256) nr += vs[i] * vs[i];
    nr = sqrt(block_sum(nr, red));
    for (int i = tid; i < n; i += 256) Xw[i] = vs[i] / nr;
    __syncthreads();
    theta[want] = th;
  }
  // outputs: eigenvalues of M ascending (= -theta), vectors with the largest-|.| component positive, breakpoint vector
  double sgn[2] = {1.0, 1.0};
  for (int q = 0; q < nev; ++q) {
    const double* Xq = q ? X1 : X0;
    double bv = 0.0; int bi = n;
    for (int i = tid; i < n; i += 256) { const double av = fabs(Xq[i]); if (av > bv) { bv = av; bi = i; } }
    // block arg-max (first index on ties)
    const double bmax = block_max(bv, red);
    int cand = (bv == bmax) ? bi : n;
    __syncthreads();
    if (tid == 0) flag[2] = n;
    __syncthreads();
    atomicMin(&flag[2], cand);
    __syncthreads();
    sgn[q] = (Xq[flag[2]] < 0.0) ? -1.0 : 1.0;
    __syncthreads();
  }
  const double l0 = -theta[0], l1 = (nev == 2) ? -theta[1] : 0.0;
  double w0 = 1.0, w1 = 0.0;
  if (nev == 2 && l1 < -1e-10) { const double nr = sqrt(l0 * l0 + l1 * l1); w0 = fabs(l0) / nr; w1 = fabs(l1) / nr; }   // OMC.jl:2471-2473
  for (int i = tid; i < n; i += 256) {
    const double v0 = sgn[0] * X0[i], v1 = (nev == 2) ? sgn[1] * X1[i] : 0.0;
    vec_out[(size_t)node * n * nev + i] = v0;
    if (nev == 2) vec_out[(size_t)node * n * nev + n + i] = v1;
    bp_out[(size_t)node * n + i] = w0 * v0 + w1 * v1;
  }
  if (tid == 0) {
    lam_out[(size_t)node * nev] = l0;
    if (nev == 2) lam_out[(size_t)node * nev + 1] = l1;
    feas_out[node] = (l0 >= -1e-6) ? 1 : 0;        // OMC.jl:1274-1276
  }
}

}  // namespace omcbig
