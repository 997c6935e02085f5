// omc_relax.cuh -- the fused per-node relaxation kernel (replaces the body of
// matrix_completion_SDP_relaxation, /root/reference/src/OptimalMatrixCompletion.jl:1431-1943).
//
// One persistent CTA per SM pulls open nodes from an atomic queue and runs the whole conic ADMM of a
// node (COSMO form) without leaving the SM: the three PSD-cone projections
// of every iteration are symmetric eigendecompositions done in shared memory (DMMA pre-rotation by
// the previous iteration's eigenvectors + parallel Jacobi sweeps + DMMA reconstruction of the smaller
// spectral side); the ADMM state of the node lives in an L2-resident global record.
#pragma once
#include "omc_device.cuh"
#include "../../include/omc_b200.h"

namespace omc {

// Layout (in doubles) of one ADMM state record: w = (X, Y, T, U), (s_b, mu_b) for the three PSD
// blocks, the box rows and the cut rows.  Symmetric matrices: row-major N x N, LOWER triangle valid.
struct StateLayout {
  int n, m, k, Lcap;
  int N1, N2, N3;
  size_t X, Y, T, U, s1, m1, s2, m2, s3, m3, s5, m5, sv, mv, sg, mg, scal, total;
};
__host__ __device__ inline StateLayout make_state_layout(int n, int m, int k, int Lcap) {
  StateLayout L;
  L.n = n; L.m = m; L.k = k; L.Lcap = Lcap;
  L.N1 = n + m; L.N2 = n + k; L.N3 = n;
  size_t o = 0;
  L.X = o; o += (size_t)n * m;
  L.Y = o; o += (size_t)n * n;
  L.T = o; o += (size_t)m * m;
  L.U = o; o += (size_t)n * k;
  L.s1 = o; o += (size_t)L.N1 * L.N1;
  L.m1 = o; o += (size_t)L.N1 * L.N1;
  L.s2 = o; o += (size_t)L.N2 * L.N2;
  L.m2 = o; o += (size_t)L.N2 * L.N2;
  L.s3 = o; o += (size_t)n * n;
  L.m3 = o; o += (size_t)n * n;
  L.s5 = o; o += (size_t)n * k;
  L.m5 = o; o += (size_t)n * k;
  L.sv = o; o += (size_t)Lcap * k;
  L.mv = o; o += (size_t)Lcap * k;
  L.sg = o; o += (size_t)Lcap;
  L.mg = o; o += (size_t)Lcap;
  L.scal = o; o += 8;  // s4, m4, rho, ncuts, iters, ...
  L.total = (o + 1) & ~(size_t)1;
  return L;
}

// per-CTA scratch (doubles): w~ (same shape as w), eigenvector bases of the 3 blocks (shared-memory
// images, NP x ld), Gram matrix of the dense rows and the Woodbury inverse, and the working state.
// Largest PSD block that is diagonalised in shared memory (two NP x ld FP64 buffers must fit in 227 KB):
// blocks with NP > OMC_SMEM_NP_MAX run through the same device functions on an L2-resident global buffer.
#define OMC_SMEM_NP_MAX 104
__host__ __device__ inline Geo smem_geo(int N1, int N2, int N3) {
  Geo best = make_geo(1);
  const int Ns[3] = {N1, N2, N3};
  for (int b = 0; b < 3; ++b) {
    const Geo g = make_geo(Ns[b]);
    if (g.NP <= OMC_SMEM_NP_MAX && g.NP > best.NP) best = g;
  }
  return best;
}

struct ScratchLayout {
  size_t wt, Q1, Q2, Q3, G, Minv, state, big0, total;
};
__host__ __device__ inline ScratchLayout make_scratch_layout(const StateLayout& S, int rmax) {
  ScratchLayout C;
  size_t o = 0;
  C.wt = o; o += (size_t)S.n * S.m + (size_t)S.n * S.n + (size_t)S.m * S.m + (size_t)S.n * S.k;
  o = (o + 1) & ~(size_t)1;
  Geo g1 = make_geo(S.N1), g2 = make_geo(S.N2), g3 = make_geo(S.N3);
  C.Q1 = o; o += (size_t)g1.NP * g1.ld;
  C.Q2 = o; o += (size_t)g2.NP * g2.ld;
  C.Q3 = o; o += (size_t)g3.NP * g3.ld;
  C.G = o; o += (size_t)rmax * rmax;
  C.Minv = o; o += (size_t)rmax * rmax;
  o = (o + 1) & ~(size_t)1;
  C.state = o; o += S.total;
  o = (o + 1) & ~(size_t)1;
  C.big0 = o;
  if (g1.NP > OMC_SMEM_NP_MAX) o += (size_t)g1.NP * g1.ld;   // working matrix of a block too large for shared memory
  C.total = (o + 1) & ~(size_t)1;
  return C;
}

struct RelaxArgs {
  int n, m, k, cut_type;
  double gamma, a, sa, cT, c0;
  const double* A;   // n*m col-major
  const double* Mk;  // n*m col-major, 0/1
  const double* pool_x;     // [cap][n]
  const double* pool_vhat;  // [cap][k]
  int B;
  const int* node_cut_ptr;
  const int* node_cut_ids;
  const uint8_t* node_cut_dirs;
  const int* warm_ids;
  const int* save_ids;
  double* pool_state;  // [cap_state][SL.total]
  double* scratch;     // [grid][SC.total]
  int* queue;
  int* status;
  double* objective;
  double* lower_bound;
  int* iters;
  double* res;
  double* outX;
  double* outY;
  double* outU;
  double* outT;
  double* prof;  // [B][16]: [8..14] = primal residual components at the last check; cycles in phase1+2, build V, gemm, jacobi, reconstruct, residual; sweeps; iterations
  omc_relax_opts o;
  int Lcap, rmax;
  StateLayout SL;
  ScratchLayout SC;
};

// (lb, ub, alpha, beta) of one column of one cut in ORIGINAL units (SURVEY.md appendix B,
// OMC.jl:1581-1676); h = vhat_j, dir = index into the label list of the cut type.
__device__ __forceinline__ void cut_coeffs(int type, int dir, double h, int fix3, double& lb, double& ub, double& al,
                                           double& be) {
  const double a = fabs(h);
  if (type == OMC_CUT_LINEAR) {
    if (dir == 0) { lb = -1.0; ub = h; al = h - 1.0; be = h; }           // OMC.jl:1582-1591
    else          { lb = h; ub = 1.0; al = h + 1.0; be = -h; }           // OMC.jl:1592-1601
  } else if (type == OMC_CUT_LINEAR2) {
    if (dir == 0)      { lb = -1.0; ub = -a; al = -(1.0 + a); be = -a; } // OMC.jl:1604-1613
    else if (dir == 1) { lb = -a; ub = a; al = 0.0; be = h * h; }        // OMC.jl:1614-1623
    else               { lb = a; ub = 1.0; al = 1.0 + a; be = -a; }      // OMC.jl:1624-1633
  } else {
    if (dir == 0)      { lb = -1.0; ub = -a; al = -(1.0 + a); be = -a; } // OMC.jl:1636-1645
    else if (dir == 1) { lb = -a; ub = 0.0; al = -a; be = 0.0; }         // OMC.jl:1646-1655
    else if (dir == 2) { lb = 0.0; ub = a; al = a; be = 0.0; }           // OMC.jl:1656-1665
    else if (fix3)     { lb = a; ub = 1.0; al = 1.0 + a; be = -a; }
    else               { lb = a; ub = 1.0; al = a; be = 0.0; }           // OMC.jl:1666-1675 (quirk Q1)
  }
}

// Everything a node's iteration needs, resolved to pointers once per node.
struct NodeCtx {
  int n, m, k, N1, N2, N3, L, r;
  double a, sa, cT, ktr, alpha, sigma, rho;
  const double* A;
  const double* Mk;
  double *X, *Y, *T, *U;      // w
  double *Xt, *Yt, *Tt, *Ut;  // w~
  double *s1, *m1, *s2, *m2, *s3, *m3, *s5, *m5, *sv, *mv, *sg, *mg, *scal;
  double *G, *Minv;
  // shared
  const double* const* cx;  // [L] pointers to cut vectors (global)
  double *lb, *ub, *al, *be;  // [L*k], [L*k], [L*k], [L]   (scaled)
  double *rhs, *cw, *gc;      // [r]
};

// z_b(r,c) = (b - A w~)_b at a lower-triangle position, and v = alpha z + (1-alpha) s + mu/rho
__device__ __forceinline__ double z_entry(const NodeCtx& c, int b, int r, int col) {
  const int n = c.n;
  if (b == 0) {
    if (r < n) return c.Yt[(size_t)r * n + col];
    if (col < n) return c.Xt[(size_t)col + (size_t)n * (r - n)];
    return c.Tt[(size_t)(r - n) * c.m + (col - n)];
  } else if (b == 1) {
    if (r < n) return c.Yt[(size_t)r * n + col];
    if (col < n) return c.Ut[(size_t)col + (size_t)n * (r - n)];
    return (r == col) ? 1.0 : 0.0;
  }
  return ((r == col) ? c.a : 0.0) - c.Yt[(size_t)r * n + col];
}
__device__ __forceinline__ double w_entry(const NodeCtx& c, int b, int r, int col) {  // same with w instead of w~
  const int n = c.n;
  if (b == 0) {
    if (r < n) return c.Y[(size_t)r * n + col];
    if (col < n) return c.X[(size_t)col + (size_t)n * (r - n)];
    return c.T[(size_t)(r - n) * c.m + (col - n)];
  } else if (b == 1) {
    if (r < n) return c.Y[(size_t)r * n + col];
    if (col < n) return c.U[(size_t)col + (size_t)n * (r - n)];
    return (r == col) ? 1.0 : 0.0;
  }
  return ((r == col) ? c.a : 0.0) - c.Y[(size_t)r * n + col];
}

// ---- dense rows R (trace row, cut rows) applied to (Yt, Ut): rhs = R [Yt; Ut] ------------------
__device__ __noinline__ void dense_rows_apply(const NodeCtx& c, const double* Yp, const double* Up, double* out,
                                              double* scratch) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int n = c.n, k = c.k, L = c.L;
  double tr = 0.0;
  for (int i = tid; i < n; i += nt) tr += Yp[(size_t)i * n + i];
  tr = block_sum(tr, scratch);
  if (tid == 0) out[0] = tr;
  // one warp per cut: xv[l,j] = x_l' Ut[:,j], q_l = x_l' Yt x_l (lower triangle, doubled off-diagonal)
  for (int l = warp; l < L; l += nw) {
    const double* x = c.cx[l];
    double q = 0.0;
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e - i * n;
      if (j <= i) {
        const double y = Yp[(size_t)i * n + j];
        q += ((i == j) ? 1.0 : 2.0) * x[i] * x[j] * y;
      }
    }
    q = warp_sum(q);
    double agg = 0.0;
    for (int j = 0; j < k; ++j) {
      double v = 0.0;
      for (int i = lane; i < n; i += 32) v += x[i] * Up[(size_t)i + (size_t)n * j];
      v = warp_sum(v);
      if (lane == 0) out[1 + l * k + j] = -v;
      agg += c.al[l * k + j] * v;
    }
    if (lane == 0) out[1 + L * k + l] = -agg + q;
  }
  __syncthreads();
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

// Minv = ((sigma + 3 rho)/rho I + G)^-1, inverted in shared memory when it fits, else in place in global
__device__ inline void build_minv(const NodeCtx& c, double* smem_work, size_t smem_cap) {
  const int tid = threadIdx.x, nt = blockDim.x, r = c.r;
  const double d = (c.sigma + 3.0 * c.rho) / c.rho;
  double* W = ((size_t)r * r <= smem_cap) ? smem_work : c.Minv;
  for (int e = tid; e < r * r; e += nt) W[e] = c.G[e] + (((e / r) == (e % r)) ? d : 0.0);
  __syncthreads();
  spd_invert(W, r);
  if (W != c.Minv)
    for (int e = tid; e < r * r; e += nt) c.Minv[e] = W[e];
  __syncthreads();
}

template <int NT, int KMAX, int MINB>
__global__ void __launch_bounds__(NT, MINB) omc_relax_kernel(const RelaxArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const int n = P.n, m = P.m, k = P.k;
  const StateLayout& SL = P.SL;
  const Geo g1 = make_geo(SL.N1), g2 = make_geo(SL.N2), g3 = make_geo(SL.N3);
  const Geo gfit = smem_geo(SL.N1, SL.N2, SL.N3);
  const size_t bufsz = (size_t)gfit.NP * gfit.ld;

  // ---- shared memory carve-up
  double* buf0 = reinterpret_cast<double*>(smem_raw);
  double* buf1 = buf0 + bufsz;
  double* lam = buf1 + bufsz;              // [NP1]
  double* wgt = lam + g1.NP;               // [NP1]
  double* jcs = wgt + g1.NP;               // [NP1/2]
  double* jsn = jcs + g1.NP / 2;           // [NP1/2]
  double* red = jsn + g1.NP / 2;           // [32]
  double* rhs = red + 32;                  // [rmax]
  double* cw = rhs + P.rmax;               // [rmax]
  double* gc = cw + P.rmax;                // [rmax]
  double* clb = gc + P.rmax;               // [Lcap*k]
  double* cub = clb + P.Lcap * k;
  double* cal = cub + P.Lcap * k;
  double* cbe = cal + P.Lcap * k;          // [Lcap]
  const double** cxp = reinterpret_cast<const double**>(cbe + P.Lcap);  // [Lcap]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(cxp + P.Lcap);           // [1]
  int* jrot = reinterpret_cast<int*>(mbar + 1);                         // [3*NP1/2]
  int* idx = jrot + 3 * (g1.NP / 2);                                    // [NP1]
  int* jskip = idx + g1.NP;                                             // [NP1] projection-mode skip flags
  int* ish = jskip + g1.NP;                                             // [8] misc ints

  double* scr = P.scratch + (size_t)blockIdx.x * P.SC.total;
  double* st = scr + P.SC.state;

  uint32_t mbar_phase = 0;
#if OMC_USE_TMA
  if (tid == 0) mbar_init(mbar, 1);
  __syncthreads();
#endif

  NodeCtx c;
  c.n = n; c.m = m; c.k = k; c.N1 = SL.N1; c.N2 = SL.N2; c.N3 = SL.N3;
  c.a = P.a; c.sa = P.sa; c.cT = P.cT; c.ktr = P.a * k; c.alpha = P.o.alpha; c.sigma = P.o.sigma;
  c.A = P.A; c.Mk = P.Mk;
  c.X = st + SL.X; c.Y = st + SL.Y; c.T = st + SL.T; c.U = st + SL.U;
  c.Xt = scr + P.SC.wt; c.Yt = c.Xt + (size_t)n * m; c.Tt = c.Yt + (size_t)n * n; c.Ut = c.Tt + (size_t)m * m;
  c.s1 = st + SL.s1; c.m1 = st + SL.m1; c.s2 = st + SL.s2; c.m2 = st + SL.m2; c.s3 = st + SL.s3; c.m3 = st + SL.m3;
  c.s5 = st + SL.s5; c.m5 = st + SL.m5; c.sv = st + SL.sv; c.mv = st + SL.mv; c.sg = st + SL.sg; c.mg = st + SL.mg;
  c.scal = st + SL.scal;
  c.G = scr + P.SC.G; c.Minv = scr + P.SC.Minv;
  c.cx = cxp; c.lb = clb; c.ub = cub; c.al = cal; c.be = cbe; c.rhs = rhs; c.cw = cw; c.gc = gc;
  double* Qg[3] = {scr + P.SC.Q1, scr + P.SC.Q2, scr + P.SC.Q3};
  double* sb[3] = {c.s1, c.s2, c.s3};
  double* mb[3] = {c.m1, c.m2, c.m3};
  const Geo gb[3] = {g1, g2, g3};

  const unsigned long long t_start = globaltimer_ns();

  for (;;) {
    // ---------------------------------------------------------------- next node from the queue
    __syncthreads();
    if (tid == 0) ish[0] = atomicAdd(P.queue, 1);
    __syncthreads();
    const int node = ish[0];
    if (node >= P.B) break;
    if (tid == 0) ish[3] = 0;
    if (P.prof)
      for (int q = tid; q < 16; q += NT) P.prof[(size_t)node * 16 + q] = 0.0;

    const int e0 = P.node_cut_ptr[node], L = P.node_cut_ptr[node + 1] - e0;
    const int r = 1 + L * (k + 1);
    c.L = L; c.r = r;
    const int warm = P.warm_ids ? P.warm_ids[node] : -1;

    // ---- cut rows (scaled): lb, ub, alpha per column, beta summed over columns
    for (int l = tid; l < L; l += NT) {
      const int cid = P.node_cut_ids[e0 + l];
      cxp[l] = P.pool_x + (size_t)cid * n;
      double bsum = 0.0;
      for (int j = 0; j < k; ++j) {
        double lb, ub, al, be;
        cut_coeffs(P.cut_type, P.node_cut_dirs[(size_t)(e0 + l) * k + j], P.pool_vhat[(size_t)cid * k + j],
                   P.o.fix_linear3_right, lb, ub, al, be);
        clb[l * k + j] = P.sa * lb; cub[l * k + j] = P.sa * ub; cal[l * k + j] = P.sa * al;
        bsum += be;
      }
      cbe[l] = P.a * bsum;
    }
    __syncthreads();

    // ---- initial state: cold (zeros, s = b) or the parent's record
    int Lw = 0;  // number of cut rows carried by the warm-start record
    if (warm >= 0) {
      const double* src = P.pool_state + (size_t)warm * SL.total;
      for (size_t e = tid; e < SL.total; e += NT) st[e] = src[e];
      __syncthreads();
      Lw = (int)c.scal[3];
      if (Lw > L) Lw = L;
      c.rho = c.scal[2];
    } else {
      for (size_t e = tid; e < SL.total; e += NT) st[e] = 0.0;
      __syncthreads();
      for (int i = tid; i < k; i += NT) c.s2[(size_t)(n + i) * SL.N2 + (n + i)] = 1.0;
      for (int i = tid; i < n; i += NT) c.s3[(size_t)i * n + i] = P.a;
      if (tid == 0) c.scal[0] = c.ktr;
      c.rho = P.o.rho0;
    }
    __syncthreads();
    // rows of the cuts this node adds on top of the warm-start record: s = clipped row value, mu = 0
    for (int l = Lw + warp; l < L; l += NW) {
      const double* x = cxp[l];
      double q = 0.0;
      for (int e = lane; e < n * n; e += 32) {
        const int i = e / n, j = e - i * n;
        if (j <= i) q += ((i == j) ? 1.0 : 2.0) * x[i] * x[j] * c.Y[(size_t)i * n + j];
      }
      q = warp_sum(q);
      double agg = cbe[l];
      for (int j = 0; j < k; ++j) {
        double v = 0.0;
        for (int i = lane; i < n; i += 32) v += x[i] * c.U[(size_t)i + (size_t)n * j];
        v = warp_sum(v);
        agg += cal[l * k + j] * v;
        if (lane == 0) {
          c.sv[l * k + j] = fmin(fmax(v, clb[l * k + j]), cub[l * k + j]);
          c.mv[l * k + j] = 0.0;
        }
      }
      if (lane == 0) {
        c.sg[l] = fmax(agg - q, 0.0);
        c.mg[l] = 0.0;
      }
    }
    __syncthreads();

    // ---- Gram matrix of the dense rows: G = R R'  (row 0 trace, 1+l*k+j the v rows, 1+L*k+l the aggregated rows)
    {
      // C[l][l'] = x_l . x_l' staged in buf0 (L*L <= bufsz is guaranteed by the host)
      double* C = buf0;
      for (int e = warp; e < L * L; e += NW) {
        const int l1 = e / L, l2 = e - l1 * L;
        double d = 0.0;
        for (int i = lane; i < n; i += 32) d += cxp[l1][i] * cxp[l2][i];
        d = warp_sum(d);
        if (lane == 0) C[e] = d;
      }
      __syncthreads();
      for (int e = tid; e < r * r; e += NT) {
        const int i = e / r, j = e - i * r;
        double v = 0.0;
        // classify rows
        const int ti = (i == 0) ? 0 : (i < 1 + L * k ? 1 : 2);
        const int tj = (j == 0) ? 0 : (j < 1 + L * k ? 1 : 2);
        const int li = (ti == 1) ? (i - 1) / k : (ti == 2 ? i - 1 - L * k : 0);
        const int lj = (tj == 1) ? (j - 1) / k : (tj == 2 ? j - 1 - L * k : 0);
        const int ci = (ti == 1) ? (i - 1) % k : 0, cj = (tj == 1) ? (j - 1) % k : 0;
        if (ti == 0 && tj == 0) v = (double)n;
        else if (ti == 0 && tj == 2) v = C[lj * L + lj];
        else if (ti == 2 && tj == 0) v = C[li * L + li];
        else if (ti == 1 && tj == 1) v = (ci == cj) ? C[li * L + lj] : 0.0;
        else if (ti == 1 && tj == 2) v = cal[lj * k + ci] * C[li * L + lj];
        else if (ti == 2 && tj == 1) v = cal[li * k + cj] * C[li * L + lj];
        else if (ti == 2 && tj == 2) {
          const double d = C[li * L + lj];
          double aa = 0.0;
          for (int j2 = 0; j2 < k; ++j2) aa += cal[li * k + j2] * cal[lj * k + j2];
          v = d * d + aa * d;
        }
        c.G[e] = v;
      }
      __syncthreads();
      build_minv(c, buf0, bufsz);
    }

    bool have_basis[3] = {false, false, false};
    // Eigensolver tolerance follows the ADMM residual: off(S) <= jtol ||S||_F with jtol two orders below the
    // current relative residual, inside [1e-13, jacobi_tol].
    double jtol = P.o.jacobi_tol;
    long long pc[6] = {0, 0, 0, 0, 0, 0};
    long long nsweeps = 0;
    long long tk = clock64();
#define OMC_TICK(slot)                 \
  {                                    \
    const long long now_ = clock64();  \
    pc[slot] += now_ - tk;             \
    tk = now_;                         \
  }
    int status = OMC_STATUS_ITERATION_LIMIT;
    double res_p = 1e300, res_d = 1e300, obj_p = 0.0, obj_d = -1e300, lbound = -1e300;
    int it = 0;

    for (it = 1; it <= P.o.max_iter; ++it) {
      const double rho = c.rho, sig = c.sigma, al = c.alpha;
      const double dYU = sig + 3.0 * rho, dT = sig + rho;
      const double t4 = c.scal[1] + rho * (c.ktr - c.scal[0]);
      // ------------------------------------------------ phase 1: w~ = D^-1 (sigma w - q + A'(rho (b - s) + mu))
      for (int e = tid; e < n * m; e += NT) {  // X, e = i + n j
        const int i = e % n, j = e / n;
        const size_t q1 = (size_t)(n + j) * SL.N1 + i;
        const double t1 = c.m1[q1] - rho * c.s1[q1];
        const double mk = c.Mk[e];
        c.Xt[e] = (sig * c.X[e] + mk * c.A[e] - 2.0 * t1) / (mk + sig + 2.0 * rho);
      }
      for (int e = tid; e < n * n; e += NT) {  // Y lower, e = i*n + j
        const int i = e / n, j = e - i * n;
        if (j > i) continue;
        const size_t q1 = (size_t)i * SL.N1 + j, q2 = (size_t)i * SL.N2 + j;
        const double t1 = c.m1[q1] - rho * c.s1[q1];
        const double t2 = c.m2[q2] - rho * c.s2[q2];
        const double t3 = c.m3[e] + rho * (((i == j) ? c.a : 0.0) - c.s3[e]);
        double gY = -t1 - t2 + t3 + ((i == j) ? t4 : 0.0);
        for (int l = 0; l < L; ++l) {
          const double tg = c.mg[l] + rho * (cbe[l] - c.sg[l]);
          gY += tg * cxp[l][i] * cxp[l][j];
        }
        c.Yt[e] = (sig * c.Y[e] + gY) / dYU;
      }
      for (int e = tid; e < m * m; e += NT) {  // Theta lower
        const int i = e / m, j = e - i * m;
        if (j > i) continue;
        const size_t q1 = (size_t)(n + i) * SL.N1 + (n + j);
        const double t1 = c.m1[q1] - rho * c.s1[q1];
        c.Tt[e] = (sig * c.T[e] - ((i == j) ? c.cT : 0.0) - t1) / dT;
      }
      for (int e = tid; e < n * k; e += NT) {  // U, e = i + n j
        const int i = e % n, j = e / n;
        const size_t q2 = (size_t)(n + j) * SL.N2 + i;
        const double t2 = c.m2[q2] - rho * c.s2[q2];
        const double t5 = c.m5[e] - rho * c.s5[e];
        double gU = -2.0 * t2 - t5;
        for (int l = 0; l < L; ++l) {
          const double tv = c.mv[l * k + j] - rho * c.sv[l * k + j];
          const double tg = c.mg[l] + rho * (cbe[l] - c.sg[l]);
          gU -= cxp[l][i] * (tv + tg * cal[l * k + j]);
        }
        c.Ut[e] = (sig * c.U[e] + gU) / dYU;
      }
      __syncthreads();
      // ------------------------------------------------ phase 2: Woodbury correction for the dense rows
      dense_rows_apply(c, c.Yt, c.Ut, rhs, red);
      for (int i = tid; i < r; i += NT) {
        double v = 0.0;
        for (int j = 0; j < r; ++j) v += c.Minv[(size_t)i * r + j] * rhs[j];
        cw[i] = v;
      }
      __syncthreads();
      for (int i = tid; i < r; i += NT) {  // gc = G cw  ->  R w~(corrected) = rhs - gc
        double v = 0.0;
        for (int j = 0; j < r; ++j) v += c.G[(size_t)i * r + j] * cw[j];
        gc[i] = v;
      }
      // w~ -= R' cw ; then w <- alpha w~ + (1-alpha) w
      for (int e = tid; e < n * n; e += NT) {
        const int i = e / n, j = e - i * n;
        if (j > i) continue;
        double corr = (i == j) ? cw[0] : 0.0;
        for (int l = 0; l < L; ++l) corr += cw[1 + L * k + l] * cxp[l][i] * cxp[l][j];
        const double yt = c.Yt[e] - corr;
        c.Yt[e] = yt;
        c.Y[e] = al * yt + (1.0 - al) * c.Y[e];
      }
      for (int e = tid; e < n * k; e += NT) {
        const int i = e % n, j = e / n;
        double corr = 0.0;
        for (int l = 0; l < L; ++l) corr -= cxp[l][i] * (cw[1 + l * k + j] + cw[1 + L * k + l] * cal[l * k + j]);
        const double ut = c.Ut[e] - corr;
        c.Ut[e] = ut;
        c.U[e] = al * ut + (1.0 - al) * c.U[e];
        // box rows
        const double v5 = al * ut + (1.0 - al) * c.s5[e] + c.m5[e] / rho;
        const double lo5 = (i >= n - k + j) ? 0.0 : -c.sa;
        const double s5n = fmin(fmax(v5, lo5), c.sa);
        c.s5[e] = s5n;
        c.m5[e] = rho * (v5 - s5n);
      }
      for (int e = tid; e < n * m; e += NT) c.X[e] = al * c.Xt[e] + (1.0 - al) * c.X[e];
      for (int e = tid; e < m * m; e += NT) {
        const int i = e / m, j = e - i * m;
        if (j <= i) c.T[e] = al * c.Tt[e] + (1.0 - al) * c.T[e];
      }
      __syncthreads();
      // scalar rows: trace, cut v rows, cut aggregated rows  (z = b - R w~)
      if (tid == 0) {
        const double z4 = c.ktr - (rhs[0] - gc[0]);
        const double v4 = al * z4 + (1.0 - al) * c.scal[0] + c.scal[1] / rho;
        const double s4n = fmax(v4, 0.0);
        c.scal[0] = s4n;
        c.scal[1] = rho * (v4 - s4n);
      }
      for (int e = tid; e < L * k; e += NT) {
        const double zv = -(rhs[1 + e] - gc[1 + e]);
        const double vv = al * zv + (1.0 - al) * c.sv[e] + c.mv[e] / rho;
        const double sn_ = fmin(fmax(vv, clb[e]), cub[e]);
        c.sv[e] = sn_;
        c.mv[e] = rho * (vv - sn_);
      }
      for (int l = tid; l < L; l += NT) {
        const double zg = cbe[l] - (rhs[1 + L * k + l] - gc[1 + L * k + l]);
        const double vg = al * zg + (1.0 - al) * c.sg[l] + c.mg[l] / rho;
        const double sn_ = fmax(vg, 0.0);
        c.sg[l] = sn_;
        c.mg[l] = rho * (vg - sn_);
      }
      __syncthreads();

      OMC_TICK(0)
      // ------------------------------------------------ phase 3: the three PSD projections
      const bool reortho = (P.o.reortho_every > 0) && (it % P.o.reortho_every == 0);
      for (int b = 0; b < 3; ++b) {
        const Geo g = gb[b];
        const int N = g.N, NP = g.NP, ld = g.ld;
        const uint32_t qbytes = (uint32_t)((size_t)NP * ld * sizeof(double));
        const bool warmQ = have_basis[b] && !reortho;
        const bool fits = (size_t)NP * ld <= bufsz;             // else: L2-resident working buffers
        double* B0 = fits ? buf0 : (scr + P.SC.big0);
        double* B1 = fits ? buf1 : Qg[b];
        // start fetching the previous eigenvector basis while V is assembled
#if OMC_USE_TMA
        if (fits && warmQ && tid == 0) {
          fence_proxy_async();
          mbar_expect_tx(mbar, qbytes);
          bulk_g2s(B1, Qg[b], qbytes, mbar);
        }
#endif
        // V = alpha z + (1-alpha) s + mu/rho, lower triangle computed, both triangles stored
        const double* sB = sb[b];
        const double* mB = mb[b];
        const double irho = 1.0 / rho;
        for (int e = tid; e < NP * NP; e += NT) {
          const int rr = e / NP, cc = e - rr * NP;
          if (cc > rr) continue;
          double v = 0.0;
          if (rr < N) {
            const size_t q = (size_t)rr * N + cc;
            v = al * z_entry(c, b, rr, cc) + (1.0 - al) * sB[q] + mB[q] * irho;
          }
          B0[(size_t)rr * ld + cc] = v;
          B0[(size_t)cc * ld + rr] = v;
        }
        OMC_TICK(1)
        if (warmQ) {
          if (fits) {
#if OMC_USE_TMA
            mbar_wait(mbar, mbar_phase);
            mbar_phase ^= 1;
#else
            for (int e = tid; e < NP * ld; e += NT) B1[e] = Qg[b][e];
#endif
          }
          __syncthreads();
          gemm_rows_inplace<KMAX>(B0, B1, NP, ld);  // W = V Q
          gemm_cols_inplace<KMAX>(B0, B1, NP, ld);  // S = Q' W
        } else {
          for (int e = tid; e < NP * NP; e += NT) {
            const int rr = e / NP, cc = e - rr * NP;
            B1[(size_t)rr * ld + cc] = (rr == cc) ? 1.0 : 0.0;
          }
          __syncthreads();
        }
        OMC_TICK(2)
        nsweeps += jacobi_sym(B0, B1, NP, ld, jtol, 40, jcs, jsn, jrot, red, 1, jskip, &ish[4],
                              P.prof ? (P.prof + (size_t)node * 16 + 8 + 3 * (b == 0 ? 0 : 1)) : nullptr);
        have_basis[b] = true;
        OMC_TICK(3)
        // eigenvalues, the smaller spectral side, compacted index list (warp 0)
        for (int i = tid; i < NP; i += NT) lam[i] = B0[(size_t)i * ld + i];
        __syncthreads();
        if (warp == 0) {
          // side to reconstruct: the one the eigensolver fully diagonalised (the other side's safe indices were
          // skipped), or the smaller one when nothing was skipped
          int side = -ish[4];
          if (side == 0) {
            int npos = 0, nneg = 0;
            for (int base = 0; base < NP; base += 32) {
              const int i = base + lane;
              const double l_ = (i < NP) ? lam[i] : 0.0;
              npos += __popc(__ballot_sync(0xffffffffu, l_ > 0.0));
              nneg += __popc(__ballot_sync(0xffffffffu, l_ < 0.0));
            }
            side = (npos <= nneg) ? 1 : -1;
          }
          int cnt = 0;
          for (int base = 0; base < NP; base += 32) {
            const int i = base + lane;
            const double l_ = (i < NP) ? lam[i] : 0.0;
            const bool pred = (i < NP) && !jskip[i < NP ? i : 0] && ((side > 0) ? (l_ > 0.0) : (l_ < 0.0));
            const unsigned bal = __ballot_sync(0xffffffffu, pred);
            if (pred) {
              const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
              idx[pos] = i;
              wgt[pos] = fabs(l_);
            }
            cnt += __popc(bal);
          }
          const int cpad = (cnt + 3) & ~3;
          if (lane < cpad - cnt) {
            idx[cnt + lane] = 0;
            wgt[cnt + lane] = 0.0;
          }
          if (lane == 0) {
            ish[1] = cpad;
            ish[2] = side;
          }
        }
#if OMC_USE_TMA
        // write the basis back for the next iteration (async proxy reads shared memory)
        fence_proxy_async();
#endif
        __syncthreads();
#if OMC_USE_TMA
        if (fits && tid == 0) {
          bulk_s2g(Qg[b], B1, qbytes);
          bulk_commit();
        }
#else
        if (fits)
          for (int e = tid; e < NP * ld; e += NT) Qg[b][e] = B1[e];
#endif
        // Z = sum_{i in side} |lam_i| q_i q_i' on lower tiles; s+ = Z (positive side) or V + Z (negative side)
        {
          const int cpad = ish[1], side = ish[2];
          const int T = NP >> 3, KS = cpad >> 2;
          const int g_ = lane >> 2, t_ = lane & 3;
          const int ntile = T * (T + 1) / 2;
          double* sW = sb[b];
          double* mW = mb[b];
          for (int tl = warp; tl < ntile; tl += NW) {
            // tile (rt, ct), rt >= ct, from the linear index
            int rt = (int)((sqrt(8.0 * tl + 1.0) - 1.0) * 0.5);
            while ((rt + 1) * (rt + 2) / 2 <= tl) ++rt;
            while (rt * (rt + 1) / 2 > tl) --rt;
            const int ct = tl - rt * (rt + 1) / 2;
            double c0 = 0.0, c1 = 0.0;
            const double* arow = B1 + (size_t)(rt * 8 + g_) * ld;
            const double* brow = B1 + (size_t)(ct * 8 + g_) * ld;
            for (int kk = 0; kk < KS; ++kk) {
              const int col = idx[kk * 4 + t_];
              const double a_ = arow[col] * wgt[kk * 4 + t_];
              const double b_ = brow[col];
              dmma884(c0, c1, a_, b_, c0, c1);
            }
            const int rr = rt * 8 + g_;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int cc = ct * 8 + 2 * t_ + h;
              if (rr < N && cc <= rr) {
                const size_t q = (size_t)rr * N + cc;
                const double v = al * z_entry(c, b, rr, cc) + (1.0 - al) * sW[q] + mW[q] * irho;
                const double z = h ? c1 : c0;
                const double snew = (side > 0) ? z : (v + z);
                sW[q] = snew;
                mW[q] = rho * (v - snew);
              }
            }
          }
        }
#if OMC_USE_TMA
        if (fits && tid == 0) bulk_wait_all();
#endif
        __syncthreads();
        OMC_TICK(4)
      }

      // ------------------------------------------------ phase 4: residuals / termination / rho
      if (it % P.o.check_every == 0 || it == P.o.max_iter) {
        tk = clock64();
        double rp = 0.0, rd = 0.0, np_ = 0.0, nd_ = 0.0, sxx = 0.0, sfit = 0.0;
        double rpc[7] = {0, 0, 0, 0, 0, 0, 0};  // components: psd1, psd2, psd3, trace, box, v rows, aggregated rows
        // PSD rows
        for (int b = 0; b < 3; ++b) {
          const int N = gb[b].N;
          const double* sB = sb[b];
          for (int e = tid; e < N * N; e += NT) {
            const int rr = e / N, cc = e - rr * N;
            if (cc > rr) continue;
            const double s_ = sB[e];
            rpc[b] = fmax(rpc[b], fabs(w_entry(c, b, rr, cc) - s_));
            np_ = fmax(np_, fabs(s_));
          }
        }
        // dual residual and norms, block by block of w
        for (int e = tid; e < n * m; e += NT) {  // X
          const int i = e % n, j = e / n;
          const double gX = -2.0 * c.m1[(size_t)(n + j) * SL.N1 + i];
          const double mk = c.Mk[e], x = c.X[e], aa = c.A[e];
          rd = fmax(rd, fabs(mk * (x - aa) - gX));
          nd_ = fmax(nd_, fmax(fabs(mk * x), fmax(fabs(mk * aa), fabs(gX))));
          sxx += mk * x * x;
          sfit += mk * (aa - x) * (aa - x);
        }
        for (int e = tid; e < n * n; e += NT) {  // Y
          const int i = e / n, j = e - i * n;
          if (j > i) continue;
          double gY = -c.m1[(size_t)i * SL.N1 + j] - c.m2[(size_t)i * SL.N2 + j] + c.m3[e] + ((i == j) ? c.scal[1] : 0.0);
          for (int l = 0; l < L; ++l) gY += c.mg[l] * cxp[l][i] * cxp[l][j];
          rd = fmax(rd, fabs(gY));
          nd_ = fmax(nd_, fabs(gY));
        }
        double trT = 0.0;
        for (int e = tid; e < m * m; e += NT) {  // Theta
          const int i = e / m, j = e - i * m;
          if (j > i) continue;
          const double gT = -c.m1[(size_t)(n + i) * SL.N1 + (n + j)];
          rd = fmax(rd, fabs(((i == j) ? c.cT : 0.0) - gT));
          nd_ = fmax(nd_, fabs(gT));
          if (i == j) trT += c.T[e];
        }
        double box_d = 0.0;
        for (int e = tid; e < n * k; e += NT) {  // U
          const int i = e % n, j = e / n;
          double gU = -2.0 * c.m2[(size_t)(n + j) * SL.N2 + i] - c.m5[e];
          for (int l = 0; l < L; ++l) gU -= cxp[l][i] * (c.mv[l * k + j] + c.mg[l] * cal[l * k + j]);
          rd = fmax(rd, fabs(gU));
          nd_ = fmax(nd_, fabs(gU));
          rpc[4] = fmax(rpc[4], fabs(c.U[e] - c.s5[e]));
          np_ = fmax(np_, fabs(c.s5[e]));
          const double lo5 = (i >= n - k + j) ? 0.0 : -c.sa;
          const double m5 = c.m5[e];
          box_d -= (m5 < 0.0) ? m5 * lo5 : m5 * c.sa;
        }
        // dense rows with the relaxed w: R [Y; U] needs a reduction -> reuse dense_rows_apply on (Y, U)
        __syncthreads();
        dense_rows_apply(c, c.Y, c.U, rhs, red);  // rhs = R w
        double dual_rows = 0.0;
#ifdef OMC_DEBUG_PRINT
        if (tid == 0 && L > 0 && it <= 3) {
          double xu = 0.0;
          for (int i = 0; i < n; ++i) xu += cxp[0][i] * c.U[i];
          printf("[dbg] it %d L %d cL %d rhs %g %g %g  xU_serial %g sv %g sg %g cbe %g nw %d\n", it, L, c.L, rhs[0], rhs[1], rhs[2], xu,
                 c.sv[0], c.sg[0], cbe[0], (int)(blockDim.x >> 5));
        }
#endif
        if (tid == 0) {
          rpc[3] = fabs((c.ktr - rhs[0]) - c.scal[0]);
          np_ = fmax(np_, fmax(fabs(c.scal[0]), fmax(c.ktr, c.a)));
          nd_ = fmax(nd_, c.cT);
          dual_rows += c.ktr * c.scal[1];
        }
        for (int e = tid; e < L * k; e += NT) {
          rpc[5] = fmax(rpc[5], fabs(-rhs[1 + e] - c.sv[e]));
          np_ = fmax(np_, fabs(c.sv[e]));
          const double mv_ = c.mv[e];
          dual_rows -= (mv_ < 0.0) ? mv_ * clb[e] : mv_ * cub[e];
        }
        for (int l = tid; l < L; l += NT) {
          rpc[6] = fmax(rpc[6], fabs((cbe[l] - rhs[1 + L * k + l]) - c.sg[l]));
          np_ = fmax(np_, fmax(fabs(c.sg[l]), fabs(cbe[l])));
          dual_rows += cbe[l] * c.mg[l];
        }
        for (int i = tid; i < k; i += NT) dual_rows += c.m2[(size_t)(n + i) * SL.N2 + (n + i)];
        for (int i = tid; i < n; i += NT) dual_rows += c.a * c.m3[(size_t)i * n + i];
        for (int q = 0; q < 7; ++q) {
          rpc[q] = block_max(rpc[q], red);
          rp = fmax(rp, rpc[q]);
        }
        rd = block_max(rd, red);
        np_ = block_max(np_, red);
        nd_ = block_max(nd_, red);
        sxx = block_sum(sxx, red);
        sfit = block_sum(sfit, red);
        trT = block_sum(trT, red);
        const double dsum = block_sum(dual_rows + box_d, red);
        res_p = rp;
        res_d = rd;
        obj_p = 0.5 * sfit + c.cT * trT;
        obj_d = -0.5 * sxx + P.c0 + dsum;
        // certified lower bound: p* >= obj_d - ||r_d||_inf * ||w*||_1, with ||w*||_1 bounded through
        // tr Y <= a k, tr Theta~ <= UB / cT, |X_ij| <= sqrt(Y_ii Theta_jj), |U| <= sqrt(a)
        {
          const double ub = (P.o.cutoff < 1e299) ? P.o.cutoff : (2.0 * fmax(fabs(obj_p), fabs(obj_d)) + 1.0);
          const double trTb = ub / c.cT;
          const double w1 = (double)n * c.ktr + sqrt((double)n * m * c.ktr * trTb) + (double)m * trTb +
                            (double)n * k * c.sa;
          lbound = obj_d - rd * w1;
        }
        bool stop = false;
        if (rp <= P.o.eps_abs + P.o.eps_rel * np_ && rd <= P.o.eps_abs + P.o.eps_rel * nd_) {
          status = OMC_STATUS_OPTIMAL;
          stop = true;
        } else if (P.o.cutoff < 1e299 && lbound > P.o.cutoff) {
          status = OMC_STATUS_CUTOFF;
          stop = true;
        } else if (P.o.time_limit_s > 0.0 &&
                   (double)(globaltimer_ns() - t_start) * 1e-9 > P.o.time_limit_s) {
          if (tid == 0) ish[3] = 1;
        }
        __syncthreads();
        if (!stop && P.o.time_limit_s > 0.0) {
          if (ish[3] == 1) {
            status = OMC_STATUS_TIME_LIMIT;
            stop = true;
          }
        }
        if (stop) break;
        {
          const double rel = fmax(rp / fmax(np_, 1.0), rd / fmax(nd_, 1.0));
          jtol = fmin(P.o.jacobi_tol, fmax(1e-13, 1e-2 * rel));
        }
        if (P.o.adapt_every > 0 && it % P.o.adapt_every == 0) {
          const double ratio = sqrt((rp / fmax(np_, 1e-12)) / fmax(rd / fmax(nd_, 1e-12), 1e-30));
          if (ratio > 5.0 || ratio < 0.2) {
            c.rho = fmin(fmax(rho * ratio, 1e-6), 1e6);
            build_minv(c, buf0, bufsz);
          }
        }
      }
    }
    if (it > P.o.max_iter) it = P.o.max_iter;
    OMC_TICK(5)
    if (P.prof && tid == 0) {
      for (int q = 0; q < 6; ++q) P.prof[(size_t)node * 16 + q] = (double)pc[q];
      P.prof[(size_t)node * 16 + 6] = (double)nsweeps;
      P.prof[(size_t)node * 16 + 7] = (double)it;
    }

    // ---------------------------------------------------------------- outputs (original units)
    if (tid == 0) {
      P.status[node] = status;
      P.objective[node] = (status == OMC_STATUS_CUTOFF) ? lbound : obj_p;
      P.lower_bound[node] = lbound;
      P.iters[node] = it;
      P.res[2 * node] = res_p;
      P.res[2 * node + 1] = res_d;
      c.scal[2] = c.rho;
      c.scal[3] = (double)L;
      c.scal[4] = (double)it;
    }
    if (P.outX)
      for (int e = tid; e < n * m; e += NT) P.outX[(size_t)node * n * m + e] = c.X[e];
    if (P.outY)
      for (int e = tid; e < n * n; e += NT) {
        const int i = e % n, j = e / n;  // output column-major (i + n j); state lower-triangle row-major
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        P.outY[(size_t)node * n * n + e] = c.Y[(size_t)hi * n + lo] / c.a;
      }
    if (P.outU)
      for (int e = tid; e < n * k; e += NT) P.outU[(size_t)node * n * k + e] = c.U[e] / c.sa;
    if (P.outT)
      for (int e = tid; e < m * m; e += NT) {
        const int i = e % m, j = e / m;
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        P.outT[(size_t)node * m * m + e] = c.T[(size_t)hi * m + lo] * c.a;
      }
    __syncthreads();
    const int save = P.save_ids ? P.save_ids[node] : -1;
    if (save >= 0) {
      double* dst = P.pool_state + (size_t)save * SL.total;
      for (size_t e = tid; e < SL.total; e += NT) dst[e] = st[e];
    }
  }
}

// shared memory bytes the kernel carves up (must mirror the carve-up above)
inline size_t relax_smem_bytes(int n, int m, int k, int Lcap, int rmax) {
  Geo g1 = make_geo(n + m);
  Geo gf = smem_geo(n + m, n + k, n);
  size_t d = 0;
  d += 2 * (size_t)gf.NP * gf.ld;        // buf0, buf1
  d += 2 * (size_t)g1.NP;                // lam, wgt
  d += 2 * (size_t)(g1.NP / 2);          // jcs, jsn
  d += 32;                               // red
  d += 3 * (size_t)rmax;                 // rhs, cw, gc
  d += 3 * (size_t)Lcap * k + Lcap;      // clb, cub, cal, cbe
  d += (size_t)Lcap;                     // cxp (pointers, 8 bytes)
  d += 1;                                // mbar
  size_t bytes = d * 8;
  bytes += sizeof(int) * (3 * ((size_t)g1.NP / 2) + 2 * (size_t)g1.NP + 8);
  return (bytes + 127) & ~(size_t)127;
}

}  // namespace omc
