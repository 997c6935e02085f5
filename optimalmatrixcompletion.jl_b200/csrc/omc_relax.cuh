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
#include "omc_lowrank.cuh"
#include "../../include/omc_b200.h"

// Infeasibility by bound is part of the default build since round 2: a node whose certified lower bound
// exceeds 1/2 ||P_Omega(A)||^2 cannot be feasible and is returned as OMC_STATUS_INFEASIBLE (OMC.jl:1921-1935).  It adds no
// pass to the kernel (one more scalar bound at each residual check); -DOMC_NO_INFEASIBLE_BY_BOUND removes it.
#if !defined(OMC_NO_INFEASIBLE_BY_BOUND) && !defined(OMC_INFEASIBLE_BY_BOUND)
#define OMC_INFEASIBLE_BY_BOUND 1
#endif

namespace omc {

// Layout (in doubles) of one ADMM state record: w = (X, Y, T, U), (s_b, mu_b) for the three PSD
// blocks, the box rows and the cut rows.  Symmetric matrices: row-major N x N, LOWER triangle valid.
struct StateLayout {
  int n, m, k, Lcap;
  int N1, N2, N3;
  size_t X, Y, T, U, s1, m1, s2, m2, s3, m3, s5, m5, sv, mv, sg, mg, scal, z1, z2, z3, total;
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
  L.scal = o; o += 8;  // s4, m4, rho, ncuts, iters, tracker flags (lr_mode_bits, lr_neg_bits, lr_p_pack)
  // tracked minority-side bases of the three PSD blocks (NP x 16 row-major, 16-byte aligned): part of the record so that a
  // child warm-started from its parent's record starts on the tracker instead of three cold eigendecompositions
  o = (o + 1) & ~(size_t)1;
  L.z1 = o; o += (size_t)((L.N1 + 7) & ~7) * 16;
  L.z2 = o; o += (size_t)((L.N2 + 7) & ~7) * 16;
  L.z3 = o; o += (size_t)((L.N3 + 7) & ~7) * 16;
  L.total = (o + 1) & ~(size_t)1;
  return L;
}

// per-CTA scratch (doubles): w~ (same shape as w), eigenvector bases of the 3 blocks (shared-memory
// images, NP x ld), Gram matrix of the dense rows and the Woodbury inverse, and the working state.
// Largest PSD block that is diagonalised in shared memory (two NP x ld FP64 buffers must fit in 227 KB):
// blocks with NP > OMC_SMEM_NP_MAX run through the same device functions on an L2-resident global buffer.
#define OMC_SMEM_NP_MAX 104
#define OMC_EPS_INF 1e-6    // relative tolerance of the primal infeasibility certificate (oracle: Options.eps_inf)
#define OMC_EPS_INF_LOOSE 1e-3
#define OMC_PROF_STRIDE 32  // doubles of per-node profile counters (omc_frontier_fetch_profile)
#define OMC_XS_CAP 3072  // doubles of shared memory reserved for a node's cut vectors (L * n <= 3072 are cached)
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
#ifdef OMC_INFEASIBILITY_CERTIFICATE
  size_t wt, Q1, Q2, Q3, Z1, Z2, Z3, G, Minv, state, big0, mold, total;
#else
  size_t wt, Q1, Q2, Q3, Z1, Z2, Z3, G, Minv, state, big0, total;
#endif
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
  C.Z1 = C.Z2 = C.Z3 = 0;              // (the tracked bases live in the state record, StateLayout::z1..z3)
  o = (o + 1) & ~(size_t)1;
  C.G = o; o += ((size_t)rmax * rmax + 1) & ~(size_t)1;      // 16-byte aligned: re-read by TMA bulk copies
  C.Minv = o; o += ((size_t)rmax * rmax + 1) & ~(size_t)1;
  o = (o + 1) & ~(size_t)1;
  C.state = o; o += S.total;
  o = (o + 1) & ~(size_t)1;
  C.big0 = o;
  if (g1.NP > OMC_SMEM_NP_MAX) o += (size_t)g1.NP * g1.ld;   // working matrix of a block too large for shared memory
#ifdef OMC_INFEASIBILITY_CERTIFICATE
  o = (o + 1) & ~(size_t)1;
  // multipliers of the previous iteration (infeasibility certificate): image of the record's span [s1, scal + 8), of
  // which only the mu parts are written
  C.mold = o; o += S.scal + 8 - S.s1;
#endif
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
  int xs_cap;   // doubles of shared memory for the node's cut vectors (L * n <= xs_cap are cached; 0 = read them from the pool)
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

// ---- dense rows R (trace row, cut rows) applied to (Y, U): out = R [Y; U] -------------------------------------
// Ys: FULL symmetric n x n with leading dimension ldy in shared memory; Up: n x k column-major (shared or global).
__device__ __forceinline__ void dense_rows_apply(const NodeCtx& c, const double* Ys, int ldy, const double* Up, double* out,
                                              double* scratch) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int n = c.n, k = c.k, L = c.L;
  double tr = 0.0;
  for (int i = tid; i < n; i += nt) tr += Ys[(size_t)i * ldy + i];
  tr = block_sum(tr, scratch);
  if (tid == 0) out[0] = tr;
  // one warp per cut: xv[l,j] = x_l' U[:,j], q_l = x_l' Y x_l (lanes own rows of Y)
  for (int l = warp; l < L; l += nw) {
    const double* x = c.cx[l];
    double q = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double* yr = Ys + (size_t)i * ldy;
      double t0 = 0.0, t1 = 0.0;
      int j = 0;
      for (; j + 1 < n; j += 2) { t0 += yr[j] * x[j]; t1 += yr[j + 1] * x[j + 1]; }
      if (j < n) t0 += yr[j] * x[j];
      q += x[i] * (t0 + t1);
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
#if OMC_USE_TMA
  fence_proxy_async_global();  // G and Minv are re-read through TMA bulk copies every iteration
#endif
  __syncthreads();
}

// doubles of the second shared-memory region: the eigenvector matrix of the full solver, or the three panels and the
// small matrices of the low-rank projection
template <int PM>
__host__ __device__ inline size_t region1_doubles(const Geo& gfit, int NPmax) {
  const size_t full = (size_t)gfit.NP * gfit.ld;
  const size_t lr = 3 * (size_t)NPmax * OMC_LR_LDZ + (sizeof(LrSmall<PM>) + 7) / 8 + 2;
  return ((full > lr ? full : lr) + 1) & ~(size_t)1;
}

// ------------------------------------------------------------------------------------------------------------------
// The kernel is split into __noinline__ phase functions that share one frame in shared memory: with everything inlined
// the register allocator kept ~200 values live across the ADMM loop and spilled into local memory inside the hot loops
// (80 % of the L2 read traffic of the first version was spill reloads).  Every phase function re-reads the few pointers it
// needs from the frame; values that change (flags, rho, mbarrier phase, counters) are written back by thread 0.
// ------------------------------------------------------------------------------------------------------------------
struct KFrame {
  double *buf0, *buf1, *lam, *wgt, *jcs, *jsn, *red, *rhs, *cw, *gc, *clb, *cub, *cal, *cbe, *tgs, *xs;
  const double** cxp;
  uint64_t* mbar;
  long long* sprof;
  int *jrot, *idx, *jskip, *ish;
  double *scr, *st, *lrZ, *lrR, *lrW;
  void* lrS;
  NodeCtx c;
  unsigned long long t_start;
  int node;
  // per-node iteration state
  unsigned have_basis_bits;   // bit b: a full eigenvector basis of block b is stored
  unsigned lr_mode_bits;      // bit b: the block's minority side is tracked by the low-rank projection
#ifdef OMC_INFEASIBILITY_CERTIFICATE
  int node_exact;             // the node looked infeasible under tracked projections: exact projections from then on
#endif
  unsigned lr_neg_bits;       // bit b: the NEGATIVE side of V is the tracked one
  unsigned lr_p_pack;         // byte b: columns of the tracked basis
  int exact_iter;             // this iteration runs exact (full) projections on every block
  int force_check;            // check residuals right after this iteration
  uint32_t mbar_phase;
  int status, it;
  double jtol, res_p, res_d, obj_p, obj_d, lbound;
  long long nsweeps, n_lr, n_full, n_idle;
  // idle tracker: a block whose minority side is EMPTY (r = 0: the projection is the identity or zero) is not re-tracked
  // while the accumulated change of V since the last tracking step, sum ||alpha (z - s)||_F, stays below half the
  // distance of its spectrum from zero (lr_margin, the top Ritz value of the tracked guard pair)
  double lr_margin[3], lr_drift[3];
  int lr_r[3];
};

#define OMC_BIT(w_, b_) (((w_) >> (b_)) & 1u)
#define OMC_SETBIT(w_, b_, v_) w_ = ((w_) & ~(1u << (b_))) | ((unsigned)((v_) ? 1u : 0u) << (b_))
#define OMC_SEL3(b_, x0, x1, x2) ((b_) == 0 ? (x0) : ((b_) == 1 ? (x1) : (x2)))
#define OMC_QG(b_) (scr + OMC_SEL3(b_, P.SC.Q1, P.SC.Q2, P.SC.Q3))
#define OMC_ZG(b_) (st + OMC_SEL3(b_, SL.z1, SL.z2, SL.z3))
#define OMC_SB(b_) (st + OMC_SEL3(b_, SL.s1, SL.s2, SL.s3))
#define OMC_MB(b_) (st + OMC_SEL3(b_, SL.m1, SL.m2, SL.m3))
#define OMC_GB(b_) make_geo(OMC_SEL3(b_, SL.N1, SL.N2, SL.N3))
#define OMC_TICK(slot)                      \
  {                                         \
    const long long now_ = clock64();       \
    if (tid == 0) sprof[slot] += now_ - tk; \
    tk = now_;                              \
  }
#define OMC_WT(slot) { const long long now_ = clock64(); if (tid == 0) sprof[16 + slot] += now_ - tw; tw = now_; }

// locals every phase function starts from (unused ones are dead code)
#define OMC_FRAME_LOCALS                                                                             \
  const int tid = threadIdx.x;                                                                       \
  const int lane = tid & 31, warp = tid >> 5;                                                        \
  constexpr int NW = NT / 32;                                                                        \
  const int n = P.n, m = P.m, k = P.k;                                                               \
  const StateLayout& SL = P.SL;                                                                      \
  const Geo g1 = make_geo(SL.N1);                                                                    \
  const Geo gfit = smem_geo(SL.N1, SL.N2, SL.N3);                                                    \
  const size_t bufsz = (size_t)gfit.NP * gfit.ld;                                                    \
  const size_t bufsz1 = region1_doubles<PM>(gfit, g1.NP);                                            \
  double* const buf0 = F.buf0; double* const buf1 = F.buf1;                                          \
  double* const lam = F.lam; double* const wgt = F.wgt; double* const jcs = F.jcs; double* const jsn = F.jsn; \
  double* const red = F.red; double* const rhs = F.rhs; double* const cw = F.cw; double* const gc = F.gc;    \
  double* const clb = F.clb; double* const cub = F.cub; double* const cal = F.cal; double* const cbe = F.cbe; \
  double* const tgs = F.tgs; double* const xs = F.xs; const double** const cxp = F.cxp;              \
  uint64_t* const mbar = F.mbar; long long* const sprof = F.sprof;                                   \
  int* const jrot = F.jrot; int* const idx = F.idx; int* const jskip = F.jskip; int* const ish = F.ish; \
  double* const scr = F.scr; double* const st = F.st;                                                \
  double* const lrZ = F.lrZ; double* const lrR = F.lrR; double* const lrW = F.lrW;                   \
  LrSmall<PM>& lrS = *reinterpret_cast<LrSmall<PM>*>(F.lrS);                                         \
  const NodeCtx& c = F.c;                                                                            \
  const int node = F.node, L = F.c.L, r = F.c.r;                                                     \
  (void)lane; (void)warp; (void)n; (void)m; (void)k; (void)g1; (void)bufsz; (void)bufsz1; (void)node; (void)L; (void)r; \
  (void)lam; (void)wgt; (void)jcs; (void)jsn; (void)red; (void)rhs; (void)cw; (void)gc; (void)clb; (void)cub; (void)cal; \
  (void)cbe; (void)tgs; (void)xs; (void)cxp; (void)mbar; (void)sprof; (void)jrot; (void)idx; (void)jskip; (void)ish; \
  (void)scr; (void)st; (void)lrZ; (void)lrR; (void)lrW; (void)lrS; (void)buf0; (void)buf1; (void)SL; (void)NW;

template <int NT, int KMAX, int PM>
__device__ __noinline__ void relax_node_setup(const RelaxArgs& P, KFrame& F) {
  OMC_FRAME_LOCALS
    const int e0 = P.node_cut_ptr[node];
    const int warm = P.warm_ids ? P.warm_ids[node] : -1;

    // ---- cut rows (scaled): lb, ub, alpha per column, beta summed over columns
    for (int l = tid; l < L; l += NT) {
      const int cid = P.node_cut_ids[e0 + l];
      cxp[l] = ((size_t)L * n <= (size_t)P.xs_cap) ? (const double*)(xs + (size_t)l * n) : (P.pool_x + (size_t)cid * n);
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

    if ((size_t)L * n <= (size_t)P.xs_cap) {
      for (int e = tid; e < L * n; e += NT) {
        const int l = e / n, i = e - l * n;
        xs[e] = P.pool_x[(size_t)P.node_cut_ids[e0 + l] * n + i];
      }
      __syncthreads();
    }
    // ---- initial state: cold (zeros, s = b) or the parent's record
    int Lw = 0;  // number of cut rows carried by the warm-start record
    double rho_init = P.o.rho0;
    if (warm >= 0) {
      const double* src = P.pool_state + (size_t)warm * SL.total;
      for (size_t e = tid; e < SL.total; e += NT) st[e] = src[e];
      __syncthreads();
      Lw = (int)c.scal[3];
      if (Lw > L) Lw = L;
      rho_init = c.scal[2];
    } else {
      for (size_t e = tid; e < SL.total; e += NT) st[e] = 0.0;
      __syncthreads();
      for (int i = tid; i < k; i += NT) c.s2[(size_t)(n + i) * SL.N2 + (n + i)] = 1.0;
      for (int i = tid; i < n; i += NT) c.s3[(size_t)i * n + i] = P.a;
      if (tid == 0) c.scal[0] = c.ktr;
      rho_init = P.o.rho0;
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
      // C[l][l'] = x_l . x_l' staged in buf0, or in the (not yet built) Minv slot of the scratch record when a small
      // problem carries more cuts than buf0 holds (r * r > L * L doubles there)
      double* C = ((size_t)L * L <= bufsz) ? buf0 : c.Minv;
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
      if (tid == 0) F.c.rho = rho_init;
      __syncthreads();
      build_minv(c, buf0, bufsz);
    }


  if (tid < 24) sprof[tid] = 0;   // [0..5] phases, [8..15] sub-phases of the low-rank step, [16..23] of the w-update
  if (tid == 0) {
    F.have_basis_bits = 0u; F.lr_mode_bits = 0u; F.lr_neg_bits = 0u; F.lr_p_pack = 0u;
    if (warm >= 0 && !P.o.exact_projection) {   // the parent's tracked bases came with its record
      F.lr_mode_bits = (unsigned)c.scal[5]; F.lr_neg_bits = (unsigned)c.scal[6]; F.lr_p_pack = (unsigned)c.scal[7];
    }
    F.exact_iter = 0; F.force_check = 0;
#ifdef OMC_INFEASIBILITY_CERTIFICATE
    F.node_exact = 0;
#endif
    // Eigensolver tolerance follows the ADMM residual: off(S) <= jtol ||S||_F with jtol two orders below the
    // current relative residual, inside [1e-13, jacobi_tol].
    F.jtol = P.o.jacobi_tol;
    F.nsweeps = 0; F.n_lr = 0; F.n_full = 0; F.n_idle = 0;
    for (int b = 0; b < 3; ++b) { F.lr_margin[b] = 0.0; F.lr_drift[b] = 0.0; F.lr_r[b] = -1; }
    F.status = OMC_STATUS_ITERATION_LIMIT;
    F.res_p = 1e300; F.res_d = 1e300; F.obj_p = 0.0; F.obj_d = -1e300; F.lbound = -1e300;
    F.it = 0;
  }
  __syncthreads();
}

// phase 1a: X and Theta rows of the w-update (own function: its register-staged loads must not compete with the rest)
template <int NT, int KMAX, int PM>
__device__ __forceinline__ void relax_p1_xt(const RelaxArgs& P, KFrame& F) {
  const int tid = threadIdx.x;
  const int n = P.n, m = P.m;
  const StateLayout& SL = P.SL;
  const NodeCtx& c = F.c;
  const double rho = c.rho, sig = c.sigma, al = c.alpha;
  const double dT = sig + rho;
      {
        constexpr int UB = 2;
        const int nm = n * m;
        for (int e0 = tid; e0 < nm; e0 += NT * UB) {  // X, e = i + n j : w~ and the relaxed w in one pass
          double vm[UB], vs[UB], vk[UB], vx[UB], va[UB];
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int e = e0 + u * NT;
            if (e < nm) {
              const int j = e / n, i = e - j * n;
              const size_t q1 = (size_t)(n + j) * SL.N1 + i;
              vm[u] = __ldcg(c.m1 + q1); vs[u] = __ldcg(c.s1 + q1); vk[u] = __ldg(c.Mk + e); vx[u] = __ldcg(c.X + e); va[u] = __ldg(c.A + e);
            }
          }
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int e = e0 + u * NT;
            if (e < nm) {
              const double t1 = vm[u] - rho * vs[u];
              const double xt = (sig * vx[u] + vk[u] * va[u] - 2.0 * t1) / (vk[u] + sig + 2.0 * rho);
              c.Xt[e] = xt;
              c.X[e] = al * xt + (1.0 - al) * vx[u];
            }
          }
        }
        const int nlt = m * (m + 1) / 2;
        for (int e0 = tid; e0 < nlt; e0 += NT * UB) {  // Theta, lower triangle by linear index
          double vm[UB], vs[UB], vt[UB];
          int ix[UB];
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int e = e0 + u * NT;
            if (e < nlt) {
              int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
              while ((i + 1) * (i + 2) / 2 <= e) ++i;
              while (i * (i + 1) / 2 > e) --i;
              const int j = e - i * (i + 1) / 2;
              const size_t q1 = (size_t)(n + i) * SL.N1 + (n + j);
              ix[u] = i * m + j;
              vm[u] = __ldcg(c.m1 + q1); vs[u] = __ldcg(c.s1 + q1); vt[u] = __ldcg(c.T + ix[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int e = e0 + u * NT;
            if (e < nlt) {
              const int i = ix[u] / m, j = ix[u] - i * m;
              const double t1 = vm[u] - rho * vs[u];
              const double tt = (sig * vt[u] - ((i == j) ? c.cT : 0.0) - t1) / dT;
              c.Tt[ix[u]] = tt;
              c.T[ix[u]] = al * tt + (1.0 - al) * vt[u];
            }
          }
        }
      }
}

// phase 1b: Y and U rows of the w-update into shared memory (before the dense-row correction)
template <int NT, int KMAX, int PM>
__device__ __forceinline__ void relax_p1_yu(const RelaxArgs& P, KFrame& F) {
  const int tid = threadIdx.x;
  const int n = P.n, k = P.k;
  const StateLayout& SL = P.SL;
  const NodeCtx& c = F.c;
  const int L = F.c.L;
  double* const buf0 = F.buf0;
  const double* const tgs = F.tgs;
  const double* const cal = F.cal;
  const double* const* const cxp = F.cxp;
  const double rho = c.rho, sig = c.sigma;
  const double dYU = sig + 3.0 * rho;
  const double t4 = c.scal[1] + rho * (c.ktr - c.scal[0]);
  const int ldy = n | 1;
  double* Ys = buf0;
  double* Us = buf0 + (size_t)n * ldy;
      {
        constexpr int UB = 2;
        const int nly = n * (n + 1) / 2;
        for (int e0 = tid; e0 < nly; e0 += NT * UB) {  // Y lower -> Ys (both triangles), before the dense-row correction
          double v1m[UB], v1s[UB], v2m[UB], v2s[UB], v3m[UB], v3s[UB], vy[UB];
          int ii[UB], jj[UB];
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int e = e0 + u * NT;
            if (e < nly) {
              int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
              while ((i + 1) * (i + 2) / 2 <= e) ++i;
              while (i * (i + 1) / 2 > e) --i;
              const int j = e - i * (i + 1) / 2;
              ii[u] = i; jj[u] = j;
              const size_t q1 = (size_t)i * SL.N1 + j, q2 = (size_t)i * SL.N2 + j, q3 = (size_t)i * n + j;
              v1m[u] = __ldcg(c.m1 + q1); v1s[u] = __ldcg(c.s1 + q1);
              v2m[u] = __ldcg(c.m2 + q2); v2s[u] = __ldcg(c.s2 + q2);
              v3m[u] = __ldcg(c.m3 + q3); v3s[u] = __ldcg(c.s3 + q3);
              vy[u] = __ldcg(c.Y + q3);
            }
          }
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int e = e0 + u * NT;
            if (e < nly) {
              const int i = ii[u], j = jj[u];
              const double t1 = v1m[u] - rho * v1s[u];
              const double t2 = v2m[u] - rho * v2s[u];
              const double t3 = v3m[u] + rho * (((i == j) ? c.a : 0.0) - v3s[u]);
              double gY = -t1 - t2 + t3 + ((i == j) ? t4 : 0.0);
              for (int l = 0; l < L; ++l) gY += tgs[l] * cxp[l][i] * cxp[l][j];
              const double yt = (sig * vy[u] + gY) / dYU;
              Ys[(size_t)i * ldy + j] = yt;
              Ys[(size_t)j * ldy + i] = yt;
            }
          }
        }
        for (int e = tid; e < n * k; e += NT) {  // U, e = i + n j -> Us
          const int i = e % n, j = e / n;
          const size_t q2 = (size_t)(n + j) * SL.N2 + i;
          const double t2 = c.m2[q2] - rho * c.s2[q2];
          const double t5 = c.m5[e] - rho * c.s5[e];
          double gU = -2.0 * t2 - t5;
          for (int l = 0; l < L; ++l) {
            const double tv = c.mv[l * k + j] - rho * c.sv[l * k + j];
            gU -= cxp[l][i] * (tv + tgs[l] * cal[l * k + j]);
          }
          Us[e] = (sig * c.U[e] + gU) / dYU;
        }
      }
}

template <int NT, int KMAX, int PM>
__device__ __noinline__ void relax_phase12(const RelaxArgs& P, KFrame& F) {
  OMC_FRAME_LOCALS
  uint32_t mbar_phase = F.mbar_phase;
  long long tk = clock64();
  {
      const double rho = c.rho, sig = c.sigma, al = c.alpha;
      (void)sig;
      // shared-memory views that live through phases 1-2 only: Y~ (full symmetric, n x ldy) and U~ in region 0, the
      // Woodbury inverse and the Gram matrix of the dense rows in region 1 (TMA bulk copies, overlapped with phase 1)
      const int ldy = n | 1;
      double* Ys = buf0;
      double* Us = buf0 + (size_t)n * ldy;
      const size_t r2p = ((size_t)r * r + 1) & ~(size_t)1;
      const bool ms_smem = 2 * r2p <= bufsz1;
      const double* Ms = ms_smem ? buf1 : c.Minv;
      const double* Gs = ms_smem ? buf1 + r2p : c.G;
#if OMC_USE_TMA
      if (ms_smem && tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(mbar, (uint32_t)(2 * r2p * sizeof(double)));
        bulk_g2s(buf1, c.Minv, (uint32_t)(r2p * sizeof(double)), mbar);
        bulk_g2s(buf1 + r2p, c.G, (uint32_t)(r2p * sizeof(double)), mbar);
      }
#else
      if (ms_smem)
        for (int e = tid; e < (int)r2p; e += NT) { buf1[e] = c.Minv[e]; buf1[r2p + e] = c.G[e]; }
#endif
      for (int l = tid; l < L; l += NT) tgs[l] = c.mg[l] + rho * (cbe[l] - c.sg[l]);
      __syncthreads();
      // ------------------------------------------------ phase 1: w~ = D^-1 (sigma w - q + A'(rho (b - s) + mu))
      // (loads of a batch are issued before any of its stores: the state record lives in L2)
      long long tw = clock64();
#define OMC_WT(slot) { const long long now_ = clock64(); if (tid == 0) sprof[16 + slot] += now_ - tw; tw = now_; }
      relax_p1_xt<NT, KMAX, PM>(P, F);
      OMC_WT(0)
      relax_p1_yu<NT, KMAX, PM>(P, F);
      __syncthreads();
      OMC_WT(1)
      // ------------------------------------------------ phase 2: Woodbury correction for the dense rows
      dense_rows_apply(c, Ys, ldy, Us, rhs, red);
      OMC_WT(2)
#if OMC_USE_TMA
      if (ms_smem) {
        mbar_wait(mbar, mbar_phase);
        mbar_phase ^= 1;
      }
#endif
      for (int i = warp; i < r; i += NW) {  // cw = Minv rhs : one warp per row
        double v = 0.0;
        for (int j = lane; j < r; j += 32) v += Ms[(size_t)i * r + j] * rhs[j];
        v = warp_sum(v);
        if (lane == 0) cw[i] = v;
      }
      __syncthreads();
      for (int i = warp; i < r; i += NW) {  // gc = G cw  ->  R w~(corrected) = rhs - gc
        double v = 0.0;
        for (int j = lane; j < r; j += 32) v += Gs[(size_t)i * r + j] * cw[j];
        v = warp_sum(v);
        if (lane == 0) gc[i] = v;
      }
      OMC_WT(3)
      // w~ -= R' cw ; then w <- alpha w~ + (1-alpha) w
      {
        const int nly = n * (n + 1) / 2;
        for (int e = tid; e < nly; e += NT) {
          int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
          while ((i + 1) * (i + 2) / 2 <= e) ++i;
          while (i * (i + 1) / 2 > e) --i;
          const int j = e - i * (i + 1) / 2;
          const size_t q3 = (size_t)i * n + j;
          const double yold = __ldcg(c.Y + q3);
          double corr = (i == j) ? cw[0] : 0.0;
          for (int l = 0; l < L; ++l) corr += cw[1 + L * k + l] * cxp[l][i] * cxp[l][j];
          const double yt = Ys[(size_t)i * ldy + j] - corr;
          c.Yt[q3] = yt;
          c.Y[q3] = al * yt + (1.0 - al) * yold;
        }
      }
      for (int e = tid; e < n * k; e += NT) {
        const int i = e % n, j = e / n;
        double corr = 0.0;
        for (int l = 0; l < L; ++l) corr -= cxp[l][i] * (cw[1 + l * k + j] + cw[1 + L * k + l] * cal[l * k + j]);
        const double ut = Us[e] - corr;
        c.Ut[e] = ut;
        c.U[e] = al * ut + (1.0 - al) * c.U[e];
        // box rows
        const double v5 = al * ut + (1.0 - al) * c.s5[e] + c.m5[e] / rho;
        const double lo5 = (i >= n - k + j) ? 0.0 : -c.sa;
        const double s5n = fmin(fmax(v5, lo5), c.sa);
        c.s5[e] = s5n;
        c.m5[e] = rho * (v5 - s5n);
      }
      __syncthreads();
      OMC_WT(4)
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
      if (tid == 0) F.mbar_phase = mbar_phase;
      __syncthreads();
  }
}

template <int NT, int KMAX, int PM>
__device__ __noinline__ void relax_project_block(const RelaxArgs& P, KFrame& F, const int b) {
  OMC_FRAME_LOCALS
  uint32_t mbar_phase = F.mbar_phase;
  unsigned have_basis_bits = F.have_basis_bits, lr_mode_bits = F.lr_mode_bits, lr_neg_bits = F.lr_neg_bits, lr_p_pack = F.lr_p_pack;
  const bool exact_iter = F.exact_iter != 0;
  const double jtol = F.jtol;
  const int it = F.it;
  long long nsweeps = 0, n_lr = 0, n_full = 0;
  const double rho = c.rho, al = c.alpha;
  const bool reortho = (P.o.reortho_every > 0) && (it % P.o.reortho_every == 0);
  long long tk = clock64();
      {
        const Geo g = OMC_GB(b);
        const int N = g.N, NP = g.NP, ld = g.ld;
        const uint32_t qbytes = (uint32_t)((size_t)NP * ld * sizeof(double));
        const bool fits = (size_t)NP * ld <= bufsz;             // else: L2-resident working buffers
#ifdef OMC_INFEASIBILITY_CERTIFICATE
        const bool node_exact = F.node_exact != 0;
        const bool use_lr = OMC_BIT(lr_mode_bits, b) && !exact_iter && !P.o.exact_projection && !node_exact;
#else
        const bool use_lr = OMC_BIT(lr_mode_bits, b) && !exact_iter && !P.o.exact_projection;
#endif
        const bool warmQ = OMC_BIT(have_basis_bits, b) && !reortho && !exact_iter && !use_lr;
        double* B0 = fits ? buf0 : (scr + P.SC.big0);
        double* B1 = fits ? buf1 : OMC_QG(b);
        // start fetching the previous eigenvector basis while V is assembled
#if OMC_USE_TMA
        if (fits && warmQ && tid == 0) {
          fence_proxy_async();
          mbar_expect_tx(mbar, qbytes);
          bulk_g2s(B1, OMC_QG(b), qbytes, mbar);
        }
#endif
        if (use_lr) {  // tracked basis -> panel (columns >= p are zero in the stored copy)
          const double2* src = reinterpret_cast<const double2*>(OMC_ZG(b));
          for (int e = tid; e < NP * 8; e += NT) {
            const int i = e >> 3, c2 = e & 7;
            *reinterpret_cast<double2*>(lrZ + (size_t)i * OMC_LR_LDZ + 2 * c2) = __ldcg(src + e);
          }
        }
        // V = alpha z + (1-alpha) s + mu/rho, lower triangle computed, both triangles stored
        const double* sB = OMC_SB(b);
        const double* mB = OMC_MB(b);
        const double irho = 1.0 / rho;
        double vsq = 0.0, dsq = 0.0;
        {
          constexpr int UB = 2;
          const int nl = N * (N + 1) / 2;
          for (int e0 = tid; e0 < nl; e0 += NT * UB) {
            double vz[UB], vs[UB], vm[UB];
            int r_[UB], c_[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
              const int e = e0 + u * NT;
              if (e < nl) {
                int rr = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                while ((rr + 1) * (rr + 2) / 2 <= e) ++rr;
                while (rr * (rr + 1) / 2 > e) --rr;
                const int cc = e - rr * (rr + 1) / 2;
                r_[u] = rr; c_[u] = cc;
                const size_t q = (size_t)rr * N + cc;
                vz[u] = z_entry(c, b, rr, cc);
                vs[u] = __ldcg(sB + q);
                vm[u] = __ldcg(mB + q);
              }
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
              const int e = e0 + u * NT;
              if (e < nl) {
                const double v = al * vz[u] + (1.0 - al) * vs[u] + vm[u] * irho;
                B0[(size_t)r_[u] * ld + c_[u]] = v;
                B0[(size_t)c_[u] * ld + r_[u]] = v;
                const double wgt2 = (r_[u] == c_[u]) ? 1.0 : 2.0, dz = vz[u] - vs[u];
                vsq += wgt2 * v * v;
                dsq += wgt2 * dz * dz;
              }
            }
          }
          const int npad = NP - N;   // zero padding rows / columns of the tile grid
          for (int e = tid; e < npad * NP; e += NT) {
            const int rr = N + e / NP, cc = e % NP;
            B0[(size_t)rr * ld + cc] = 0.0;
            B0[(size_t)cc * ld + rr] = 0.0;
          }
        }
        OMC_TICK(1)
        bool lr_done = false;
        double* lrOut = lrW;
        if (use_lr) {
          const double vscale = sqrt(block_sum(vsq, red));
          const double dstep = al * sqrt(block_sum(dsq, red));   // ||V - V_prev||_F = alpha ||z - s||_F
          __syncthreads();
          const bool idle = !P.o.exact_projection && F.lr_r[b] == 0 && (F.lr_drift[b] + dstep) < 0.5 * F.lr_margin[b];
          if (idle) {  // nothing can have crossed zero: s+ = V (or 0), mu+ = 0 (or rho V); the guard pair is kept as it is
            lr_done = true;
            __syncthreads();
            if (tid == 0) {
              F.lr_drift[b] += dstep;
              F.n_idle += 1;
              ish[1] = 0;
              ish[2] = OMC_BIT(lr_neg_bits, b) ? -1 : 1;
            }
            __syncthreads();
          }
          const int need_full = idle ? 0 : lowrank_step<PM>(B0, ld, !fits, N, NP, OMC_BIT(lr_neg_bits, b) ? -1.0 : 1.0, lrZ, lrR, lrW, (int)((lr_p_pack >> (8 * b)) & 0xffu), lrS, vscale, P.prof ? sprof + 8 : nullptr, &lrOut);
          if (!idle) ++n_lr;
          OMC_TICK(2)
          if (idle) {
          } else if (!need_full) {
            lr_done = true;
            const int pn = lrS.info[0], r_ = lrS.info[1];
            lr_p_pack = (lr_p_pack & ~(0xffu << (8 * b))) | ((unsigned)pn << (8 * b));
            // store the new basis (16 columns; columns >= pn zero) and set up the reconstruction
            for (int e = tid; e < NP * 16; e += NT) {
              const int i = e >> 4, j = e & 15;
              OMC_ZG(b)[e] = (j < pn) ? lrOut[(size_t)i * OMC_LR_LDZ + j] : 0.0;
            }
            if (tid < 16) {
              idx[tid] = (tid < r_) ? tid : 0;
              wgt[tid] = (tid < r_) ? lrS.th[tid] : 0.0;
            }
            if (tid == 0) {
              F.lr_r[b] = r_;
              F.lr_drift[b] = 0.0;
              F.lr_margin[b] = (r_ == 0 && pn > 0 && -lrS.th[0] > 1e-4 * vscale) ? -lrS.th[0] : 0.0;
              ish[1] = (r_ + 3) & ~3;
              ish[2] = OMC_BIT(lr_neg_bits, b) ? -1 : 1;
            }
            __syncthreads();
          } else {
            OMC_SETBIT(lr_mode_bits, b, 0);   // minority side outgrew the panel: full solve on the intact V, cold basis
            OMC_SETBIT(have_basis_bits, b, 0);
            __syncthreads();
          }
        }
        if (!lr_done) {
          ++n_full;
          const bool warm_now = warmQ && !use_lr;
          if (warm_now) {
            if (fits) {
#if OMC_USE_TMA
              mbar_wait(mbar, mbar_phase);
              mbar_phase ^= 1;
#else
              for (int e = tid; e < NP * ld; e += NT) B1[e] = OMC_QG(b)[e];
#endif
            }
            __syncthreads();
            gemm_rows_inplace<KMAX>(B0, B1, NP, ld);  // W = V Q
            gemm_cols_inplace<KMAX>(B0, B1, NP, ld);  // S = Q' W
          } else {
            __syncthreads();
            for (int e = tid; e < NP * NP; e += NT) {
              const int rr = e / NP, cc = e - rr * NP;
              B1[(size_t)rr * ld + cc] = (rr == cc) ? 1.0 : 0.0;
            }
            __syncthreads();
          }
          OMC_TICK(2)
          // a cold solve (first iteration, re-orthogonalisation, tracker fallback) runs to the tight tolerance: the loose jtol is
          // only sound for warm solves, whose error is coherent from one iteration to the next (with cold solves every
          // projection carries an independent O(jtol) error and the residual stalls at ~100 jtol: seen on 6 x 6 blocks)
#ifdef OMC_INFEASIBILITY_CERTIFICATE
          double tol_here = (exact_iter || !warm_now) ? fmin(jtol, 1e-10) : jtol;
          // nodes with cuts can be infeasible: their iterates diverge, ||V|| grows and a tolerance relative to ||V|| alone
          // would freeze the warm basis (the step falls below it), spoiling d mu = mu_it - mu_(it-1) of the certificate.
          // Resolve the step itself to 1e-4; a node already under suspicion runs at full accuracy.
          if (node_exact) {
            tol_here = 1e-13;
          } else if (L > 0 && fits) {
            const double vs_ = sqrt(block_sum(vsq, red)), ds_ = al * sqrt(block_sum(dsq, red));
            tol_here = fmin(tol_here, fmax(1e-13, 1e-4 * ds_ / fmax(vs_, 1e-300)));
          }
#else
          const double tol_here = (exact_iter || !warm_now) ? fmin(jtol, 1e-10) : jtol;
#endif
          nsweeps += jacobi_sym(B0, B1, NP, ld, tol_here, 40, jcs, jsn, jrot, red, 1, jskip, &ish[4],
                                P.prof ? (P.prof + (size_t)node * OMC_PROF_STRIDE + 8 + 3 * (b == 0 ? 0 : 1)) : nullptr);
          OMC_SETBIT(have_basis_bits, b, 1);
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
              ish[5] = cnt;
            }
          }
#if OMC_USE_TMA
          // write the basis back for the next iteration (async proxy reads shared memory)
          fence_proxy_async();
#endif
          __syncthreads();
#if OMC_USE_TMA
          if (fits && tid == 0) {
            bulk_s2g(OMC_QG(b), B1, qbytes);
            bulk_commit();
          }
#else
          if (fits)
            for (int e = tid; e < NP * ld; e += NT) OMC_QG(b)[e] = B1[e];
#endif
          // switch to the low-rank projection when the minority side (plus guard band) fits the panel: the tracked
          // basis = the minority-side eigenvectors and the OMC_LR_BUF eigenvectors next to them across zero
          // (only when the Rayleigh-Ritz space [Z R~] of 2 p directions fits the block: small blocks stay on the full solver)
#ifdef OMC_INFEASIBILITY_CERTIFICATE
          if (!P.o.exact_projection && !node_exact && ish[5] + OMC_LR_BUF <= PM && 2 * (ish[5] + OMC_LR_BUF) <= N) {
#else
          if (!P.o.exact_projection && ish[5] + OMC_LR_BUF <= PM && 2 * (ish[5] + OMC_LR_BUF) <= N) {
#endif
            const int side = ish[2], pz = ish[5] + OMC_LR_BUF;
            if (tid < 16) jrot[tid] = -1;
            __syncthreads();
            if (tid < N) {
              const double key = side * lam[tid];
              int rank = 0;
              for (int j = 0; j < N; ++j) {
                const double o = side * lam[j];
                if (o > key || (o == key && j < tid)) ++rank;
              }
              if (rank < pz) jrot[rank] = tid;
            }
            __syncthreads();
            for (int e = tid; e < NP * 16; e += NT) {
              const int i = e >> 4, j = e & 15;
              const int col = jrot[j];
              OMC_ZG(b)[e] = (col >= 0 && i < N) ? B1[(size_t)i * ld + col] : 0.0;
            }
            OMC_SETBIT(lr_mode_bits, b, 1);
            if (tid == 0) { F.lr_r[b] = -1; F.lr_margin[b] = 0.0; F.lr_drift[b] = 0.0; }
            lr_p_pack = (lr_p_pack & ~(0xffu << (8 * b))) | ((unsigned)pz << (8 * b));
            OMC_SETBIT(lr_neg_bits, b, side < 0);
          } else {
            OMC_SETBIT(lr_mode_bits, b, 0);
          }
        }
        // Z = sum_{i in side} |lam_i| q_i q_i' on lower tiles; s+ = Z (positive side) or V + Z (negative side)
        {
          const int cpad = ish[1], side = ish[2];
          const int T = NP >> 3, KS = cpad >> 2;
          const int g_ = lane >> 2, t_ = lane & 3;
          const int ntile = T * (T + 1) / 2;
          const double* QB = lr_done ? lrOut : B1;
          const int ldq = lr_done ? OMC_LR_LDZ : ld;
          double* sW = OMC_SB(b);
          double* mW = OMC_MB(b);
          for (int tl = warp; tl < ntile; tl += NW) {
            // tile (rt, ct), rt >= ct, from the linear index
            int rt = (int)((sqrt(8.0 * tl + 1.0) - 1.0) * 0.5);
            while ((rt + 1) * (rt + 2) / 2 <= tl) ++rt;
            while (rt * (rt + 1) / 2 > tl) --rt;
            const int ct = tl - rt * (rt + 1) / 2;
            double c0 = 0.0, c1 = 0.0;
            const double* arow = QB + (size_t)(rt * 8 + g_) * ldq;
            const double* brow = QB + (size_t)(ct * 8 + g_) * ldq;
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
                const double v = lr_done ? B0[(size_t)rr * ld + cc]
                                         : (al * z_entry(c, b, rr, cc) + (1.0 - al) * sW[q] + mW[q] * irho);
                const double z = h ? c1 : c0;
                const double snew = (side > 0) ? z : (v + z);
                sW[q] = snew;
                mW[q] = rho * (v - snew);
              }
            }
          }
        }
#if OMC_USE_TMA
        if (!lr_done && fits && tid == 0) bulk_wait_all();
#endif
        __syncthreads();
        OMC_TICK(4)
      }
  if (tid == 0) {
    F.mbar_phase = mbar_phase;
    F.have_basis_bits = have_basis_bits; F.lr_mode_bits = lr_mode_bits; F.lr_neg_bits = lr_neg_bits; F.lr_p_pack = lr_p_pack;
    F.nsweeps += nsweeps; F.n_lr += n_lr; F.n_full += n_full;
  }
  __syncthreads();
}

// returns true when the node is finished
template <int NT, int KMAX, int PM>
__device__ __noinline__ bool relax_phase4(const RelaxArgs& P, KFrame& F) {
  OMC_FRAME_LOCALS
  const int it = F.it;
  const unsigned lr_mode_bits = F.lr_mode_bits;
  bool exact_iter = F.exact_iter != 0, force_check = false;
  double jtol = F.jtol;
  int status = F.status;
  double res_p, res_d, obj_p, obj_d, lbound;
  const double rho = c.rho;
  const unsigned long long t_start = F.t_start;
  long long tk;
  __syncthreads();   // every thread has read the frame before thread 0 rewrites it below
      {
        tk = clock64();
        // a termination decision taken on tracked (low-rank) projections is only provisional: it is re-taken right
        // after one iteration with exact projections on every block (s in the cone and mu in its polar exactly)
#ifdef OMC_INFEASIBILITY_CERTIFICATE
        const bool node_exact = F.node_exact != 0;
        const bool provisional = !exact_iter && lr_mode_bits != 0u && !P.o.exact_projection && !node_exact;
#else
        const bool provisional = !exact_iter && lr_mode_bits != 0u && !P.o.exact_projection;
#endif
        exact_iter = false;
        force_check = false;
        double rp = 0.0, rd = 0.0, np_ = 0.0, nd_ = 0.0, sxx = 0.0, sfit = 0.0;
        double rpc[7] = {0, 0, 0, 0, 0, 0, 0};  // components: psd1, psd2, psd3, trace, box, v rows, aggregated rows
        // PSD rows
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const int N = OMC_GB(b).N;
          const double* sB = OMC_SB(b);
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
        {  // stage the relaxed Y (lower triangle in the record) as a full symmetric matrix in region 0
          const int ldy = n | 1;
          for (int e = tid; e < n * n; e += NT) {
            const int i = e / n, j = e - i * n;
            if (j <= i) {
              const double y = __ldcg(c.Y + e);
              buf0[(size_t)i * ldy + j] = y;
              buf0[(size_t)j * ldy + i] = y;
            }
          }
          __syncthreads();
          dense_rows_apply(c, buf0, ldy, c.U, rhs, red);  // rhs = R w
        }
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
#ifdef OMC_INFEASIBLE_BY_BOUND
        // Infeasibility by bound (validated in the oracle, Options.infeasible_by_bound): a feasible node has p* <= c0 = 1/2 ||P_Omega(A)||^2 (X = 0, Theta = 0 with any feasible (Y, U)), so
        // ||w*||_1 <= w1(c0) and a certified bound above c0 contradicts feasibility.  No extra pass; on the infeasible chain
        // of the tests it fires at iteration 400 where the d mu certificate needs 5 450.
        double lbound_c0;
        {
          const double trTb = P.c0 / c.cT;
          const double w1 = (double)n * c.ktr + sqrt((double)n * m * c.ktr * trTb) + (double)m * trTb + (double)n * k * c.sa;
          lbound_c0 = obj_d - rd * w1;
        }
#endif
        bool stop = false;
        if (!(fabs(obj_p) < 1e300 && fabs(obj_d) < 1e300 && rp < 1e300 && rd < 1e300)) {  // NaN / overflow (fmax drops NaN)
          status = OMC_STATUS_NUMERICAL;
          stop = true;
        } else if (rp <= P.o.eps_abs + P.o.eps_rel * np_ && rd <= P.o.eps_abs + P.o.eps_rel * nd_) {
          if (provisional && it < P.o.max_iter) {
            exact_iter = true;
            force_check = true;
          } else {
            status = OMC_STATUS_OPTIMAL;
            stop = true;
          }
        } else if (P.o.cutoff < 1e299 && lbound > P.o.cutoff) {
          if (provisional && it < P.o.max_iter) {
            exact_iter = true;
            force_check = true;
          } else {
            status = OMC_STATUS_CUTOFF;
            stop = true;
          }
#ifdef OMC_INFEASIBLE_BY_BOUND
        } else if (L > 0 && lbound_c0 > P.c0 * (1.0 + 1e-9) + 1e-12) {
          if (provisional && it < P.o.max_iter) {
            exact_iter = true;
            force_check = true;
          } else {
            status = OMC_STATUS_INFEASIBLE;
            stop = true;
          }
#endif
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
        if (stop) {
          if (tid == 0) {
            F.status = status; F.res_p = res_p; F.res_d = res_d; F.obj_p = obj_p; F.obj_d = obj_d; F.lbound = lbound;
            F.exact_iter = 0; F.force_check = 0;
          }
          __syncthreads();
          return true;
        }
#ifdef OMC_INFEASIBILITY_CERTIFICATE
        // ---- primal infeasibility certificate (oracle/relaxation.py, COSMO sec. 5.2): d = mu - mu_prev in the polar cone,
        // A'd ~ 0 and support(d) - b'd < 0.  Only for nodes with cuts (the root is always feasible) whose blocks fit the
        // shared-memory eigensolver; the eigenvalue test runs only when the cheap conditions hold.
        if (L > 0 && (size_t)g1.NP * g1.ld <= bufsz) {
          const double* mo = scr + P.SC.mold;
          const size_t base = SL.s1;
#define OMC_DM(off_) (__ldcg(st + (off_)) - __ldcg(mo + ((off_) - base)))
          double* dvs = cw;   // [L*k] d of the v rows
          double* dgs = gc;   // [L]   d of the aggregated rows
          for (int e = tid; e < L * k; e += NT) dvs[e] = OMC_DM(SL.mv + e);
          for (int l = tid; l < L; l += NT) dgs[l] = OMC_DM(SL.mg + l);
          __syncthreads();
          const double d4 = OMC_DM(SL.scal + 1);
          double nrm = fabs(d4), atn = 0.0, sup = 0.0, bdy = 0.0, cmax = d4;
          for (int e = tid; e < n * m; e += NT) {  // X
            const int i = e % n, j = e / n;
            const double d = OMC_DM(SL.m1 + (size_t)(n + j) * SL.N1 + i);
            nrm = fmax(nrm, fabs(d));
            atn = fmax(atn, 2.0 * fabs(d));
          }
          for (int e = tid; e < n * n; e += NT) {  // Y
            const int i = e / n, j = e - i * n;
            if (j > i) continue;
            const double d1 = OMC_DM(SL.m1 + (size_t)i * SL.N1 + j), d2 = OMC_DM(SL.m2 + (size_t)i * SL.N2 + j), d3 = OMC_DM(SL.m3 + e);
            double g = -d1 - d2 + d3 + ((i == j) ? d4 : 0.0);
            for (int l = 0; l < L; ++l) g += dgs[l] * cxp[l][i] * cxp[l][j];
            nrm = fmax(nrm, fmax(fabs(d1), fmax(fabs(d2), fabs(d3))));
            atn = fmax(atn, fabs(g));
            if (i == j) bdy += c.a * d3;
          }
          for (int e = tid; e < m * m; e += NT) {  // Theta
            const int i = e / m, j = e - i * m;
            if (j > i) continue;
            const double d = OMC_DM(SL.m1 + (size_t)(n + i) * SL.N1 + (n + j));
            nrm = fmax(nrm, fabs(d));
            atn = fmax(atn, fabs(d));
          }
          for (int e = tid; e < n * k; e += NT) {  // U and the box rows
            const int i = e % n, j = e / n;
            const double d2 = OMC_DM(SL.m2 + (size_t)(n + j) * SL.N2 + i), d5 = OMC_DM(SL.m5 + e);
            double g = -2.0 * d2 - d5;
            for (int l = 0; l < L; ++l) g -= cxp[l][i] * (dvs[l * k + j] + dgs[l] * cal[l * k + j]);
            nrm = fmax(nrm, fmax(fabs(d2), fabs(d5)));
            atn = fmax(atn, fabs(g));
            const double lo5 = (i >= n - k + j) ? 0.0 : -c.sa;
            sup += (d5 > 0.0) ? c.sa * d5 : lo5 * d5;
          }
          for (int e = tid; e < k * k; e += NT) {  // identity corner of the second block
            const int i = e / k, j = e - i * k;
            if (j > i) continue;
            const double d = OMC_DM(SL.m2 + (size_t)(n + i) * SL.N2 + (n + j));
            nrm = fmax(nrm, fabs(d));
            if (i == j) bdy += d;
          }
          for (int e = tid; e < L * k; e += NT) {
            const double d = dvs[e];
            nrm = fmax(nrm, fabs(d));
            sup += (d > 0.0) ? cub[e] * d : clb[e] * d;
          }
          for (int l = tid; l < L; l += NT) {
            const double d = dgs[l];
            nrm = fmax(nrm, fabs(d));
            bdy += cbe[l] * d;
            cmax = fmax(cmax, d);
          }
          if (tid == 0) bdy += c.ktr * d4;
          nrm = block_max(nrm, red);
          atn = block_max(atn, red);
          cmax = block_max(cmax, red);
          sup = block_sum(sup, red);
          bdy = block_sum(bdy, red);
          // tracked projections (and warm eigensolves at the loose tolerance) leave an error floor of ~1e-5 on A'd: a node
          // that LOOKS infeasible at a loose tolerance finishes on exact projections at full eigensolver accuracy, where the
          // strict certificate can be met (same rule in the oracle)
          if (!node_exact && nrm > 1e-14 && atn <= OMC_EPS_INF_LOOSE * nrm && cmax <= OMC_EPS_INF_LOOSE * nrm &&
              sup - bdy < -OMC_EPS_INF_LOOSE * nrm) {
            if (tid == 0) F.node_exact = 1;
          }
          const double tol = OMC_EPS_INF * nrm;
          if (nrm > 1e-14 && atn <= tol && cmax <= tol && sup - bdy < -tol) {
            bool cone_ok = true;
            for (int b = 0; b < 3 && cone_ok; ++b) {  // lambda_max(d_b) <= tol
              const Geo g = OMC_GB(b);
              const size_t mb = (b == 0) ? SL.m1 : ((b == 1) ? SL.m2 : SL.m3);
              __syncthreads();
              for (int e = tid; e < g.NP * g.NP; e += NT) {
                const int rr = e / g.NP, cc = e - rr * g.NP;
                if (cc <= rr) {
                  const double d = (rr < g.N) ? OMC_DM(mb + (size_t)rr * g.N + cc) : 0.0;
                  buf0[(size_t)rr * g.ld + cc] = d;
                  buf0[(size_t)cc * g.ld + rr] = d;
                }
                buf1[(size_t)rr * g.ld + cc] = (rr == cc) ? 1.0 : 0.0;
              }
              __syncthreads();
              jacobi_sym(buf0, buf1, g.NP, g.ld, 1e-12, 40, jcs, jsn, jrot, red);
              double lmax = -1e300;
              for (int i = tid; i < g.N; i += NT) lmax = fmax(lmax, buf0[(size_t)i * g.ld + i]);
              lmax = block_max(lmax, red);
              if (lmax > tol) cone_ok = false;
            }
            if (cone_ok) {
              if (tid == 0) {
                F.status = OMC_STATUS_INFEASIBLE; F.res_p = res_p; F.res_d = res_d; F.obj_p = obj_p; F.obj_d = obj_d; F.lbound = lbound;
                F.exact_iter = 0; F.force_check = 0;
              }
              __syncthreads();
              return true;
            }
          }
#undef OMC_DM
        }
#endif
        {
          const double rel = fmax(rp / fmax(np_, 1.0), rd / fmax(nd_, 1.0));
          jtol = fmin(P.o.jacobi_tol, fmax(1e-13, 1e-2 * rel));
        }
        if (P.o.adapt_every > 0 && it % P.o.adapt_every == 0) {
          const double ratio = sqrt((rp / fmax(np_, 1e-12)) / fmax(rd / fmax(nd_, 1e-12), 1e-30));
          if (ratio > 5.0 || ratio < 0.2) {
            if (tid == 0) F.c.rho = fmin(fmax(rho * ratio, 1e-6), 1e6);
            __syncthreads();
            build_minv(c, buf0, bufsz);
          }
        }
      }
  if (tid == 0) {
    F.status = status; F.res_p = res_p; F.res_d = res_d; F.obj_p = obj_p; F.obj_d = obj_d; F.lbound = lbound;
    F.exact_iter = exact_iter ? 1 : 0; F.force_check = force_check ? 1 : 0; F.jtol = jtol;
  }
  __syncthreads();
  return false;
}

#ifdef OMC_INFEASIBILITY_CERTIFICATE
// mu of the previous iteration, kept for the infeasibility certificate of the next check (d mu = mu_it - mu_(it-1),
// COSMO sec. 5.2 as restated in oracle/relaxation.py); called only at the top of an iteration that ends with a check.
template <int NT, int KMAX, int PM>
__device__ __noinline__ void relax_save_mu(const RelaxArgs& P, KFrame& F) {
  OMC_FRAME_LOCALS
  (void)c;
  double* mo = scr + P.SC.mold;
  const size_t base = SL.s1;
  const size_t off[6] = {SL.m1, SL.m2, SL.m3, SL.m5, SL.mv, SL.mg};
  const size_t len[6] = {(size_t)SL.N1 * SL.N1, (size_t)SL.N2 * SL.N2, (size_t)n * n, (size_t)n * k, (size_t)L * k, (size_t)L};
  for (int q = 0; q < 6; ++q)
    for (size_t e = tid; e < len[q]; e += NT) mo[off[q] - base + e] = __ldcg(st + off[q] + e);
  if (tid == 0) mo[SL.scal + 1 - base] = st[SL.scal + 1];
  __syncthreads();
}
#endif

template <int NT, int KMAX, int PM>
__device__ __noinline__ void relax_node_output(const RelaxArgs& P, KFrame& F) {
  OMC_FRAME_LOCALS
  const int it = F.it > P.o.max_iter ? P.o.max_iter : F.it;
  const int status = F.status;
  const double res_p = F.res_p, res_d = F.res_d, obj_p = F.obj_p, lbound = F.lbound;
  const long long nsweeps = F.nsweeps, n_lr = F.n_lr, n_full = F.n_full, n_idle = F.n_idle;
  long long tk = clock64();
    OMC_TICK(5)
    if (P.prof && tid == 0) {
      for (int q = 0; q < 6; ++q) P.prof[(size_t)node * OMC_PROF_STRIDE + q] = (double)sprof[q];
      P.prof[(size_t)node * OMC_PROF_STRIDE + 6] = (double)nsweeps;
      P.prof[(size_t)node * OMC_PROF_STRIDE + 7] = (double)it;
      P.prof[(size_t)node * OMC_PROF_STRIDE + 14] = (double)n_lr;
      P.prof[(size_t)node * OMC_PROF_STRIDE + 15] = (double)n_full;
      P.prof[(size_t)node * OMC_PROF_STRIDE + 13] = (double)n_idle;
      for (int q = 0; q < 8; ++q) P.prof[(size_t)node * OMC_PROF_STRIDE + 16 + q] = (double)sprof[8 + q];
      for (int q = 0; q < 8; ++q) P.prof[(size_t)node * OMC_PROF_STRIDE + 24 + q] = (double)sprof[16 + q];
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
      c.scal[5] = (double)F.lr_mode_bits; c.scal[6] = (double)F.lr_neg_bits; c.scal[7] = (double)F.lr_p_pack;
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
  __syncthreads();
}

template <int NT, int KMAX, int MINB, int PM>
__global__ void __launch_bounds__(NT, MINB) omc_relax_kernel(const __grid_constant__ RelaxArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int k = P.k;
  const StateLayout& SL = P.SL;
  const Geo g1 = make_geo(SL.N1);
  const Geo gfit = smem_geo(SL.N1, SL.N2, SL.N3);
  const size_t bufsz = (size_t)gfit.NP * gfit.ld;

  // ---- shared memory carve-up
  double* buf0 = reinterpret_cast<double*>(smem_raw);
  double* buf1 = buf0 + bufsz;
  const size_t bufsz1 = region1_doubles<PM>(gfit, g1.NP);
  // the per-index arrays hold at least 16 entries: the tracker's panel bookkeeping (idx / wgt / jrot[0..16)) is sized
  // by the panel, not by the block (blocks of 8 rows overflowed them: seen as a stalled ADMM on 3 x 3 and 4 x 4 inputs)
  const int NPI = g1.NP < 16 ? 16 : g1.NP;
  double* lam = buf1 + bufsz1;             // [NPI]
  double* wgt = lam + NPI;                 // [NPI]
  double* jcs = wgt + NPI;                 // [NPI/2]
  double* jsn = jcs + NPI / 2;             // [NPI/2]
  double* red = jsn + NPI / 2;             // [32]
  double* rhs = red + 32;                  // [rmax]
  double* cw = rhs + P.rmax;               // [rmax]
  double* gc = cw + P.rmax;                // [rmax]
  double* clb = gc + P.rmax;               // [Lcap*k]
  double* cub = clb + P.Lcap * k;
  double* cal = cub + P.Lcap * k;
  double* cbe = cal + P.Lcap * k;          // [Lcap]
  double* tgs = cbe + P.Lcap;              // [Lcap]  rho (beta - sg) + mg of the aggregated rows, refreshed every iteration
  double* xs = tgs + P.Lcap;               // [xs_cap] shared-memory copies of the node's cut vectors (when they fit)
  const double** cxp = reinterpret_cast<const double**>(xs + P.xs_cap);  // [Lcap]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(cxp + P.Lcap);           // [1]
  long long* sprof = reinterpret_cast<long long*>(mbar + 1);            // [24] cycle counters (thread 0 accumulates)
  KFrame& F = *reinterpret_cast<KFrame*>(sprof + 24);                   // the shared frame
  int* jrot = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(&F) + ((sizeof(KFrame) + 7) & ~(size_t)7));  // [3*NPI/2]
  int* idx = jrot + 3 * (NPI / 2);                                      // [NPI]
  int* jskip = idx + NPI;                                               // [NPI] projection-mode skip flags
  int* ish = jskip + NPI;                                               // [8] misc ints

  if (tid == 0) {
    double* scr = P.scratch + (size_t)blockIdx.x * P.SC.total;
    double* st = scr + P.SC.state;
    const int n = P.n, m = P.m;
    F.buf0 = buf0; F.buf1 = buf1; F.lam = lam; F.wgt = wgt; F.jcs = jcs; F.jsn = jsn; F.red = red; F.rhs = rhs; F.cw = cw; F.gc = gc;
    F.clb = clb; F.cub = cub; F.cal = cal; F.cbe = cbe; F.tgs = tgs; F.xs = xs; F.cxp = cxp; F.mbar = mbar; F.sprof = sprof;
    F.jrot = jrot; F.idx = idx; F.jskip = jskip; F.ish = ish; F.scr = scr; F.st = st;
    // low-rank projection workspace inside region 1 (panels first, then the small matrices)
    F.lrZ = buf1;
    F.lrR = F.lrZ + (size_t)g1.NP * OMC_LR_LDZ;
    F.lrW = F.lrR + (size_t)g1.NP * OMC_LR_LDZ;
    F.lrS = F.lrW + (size_t)g1.NP * OMC_LR_LDZ;
    NodeCtx& c = F.c;
    c.n = n; c.m = m; c.k = k; c.N1 = SL.N1; c.N2 = SL.N2; c.N3 = SL.N3; c.L = 0; c.r = 1;
    c.a = P.a; c.sa = P.sa; c.cT = P.cT; c.ktr = P.a * k; c.alpha = P.o.alpha; c.sigma = P.o.sigma; c.rho = P.o.rho0;
    c.A = P.A; c.Mk = P.Mk;
    c.X = st + SL.X; c.Y = st + SL.Y; c.T = st + SL.T; c.U = st + SL.U;
    c.Xt = scr + P.SC.wt; c.Yt = c.Xt + (size_t)n * m; c.Tt = c.Yt + (size_t)n * n; c.Ut = c.Tt + (size_t)m * m;
    c.s1 = st + SL.s1; c.m1 = st + SL.m1; c.s2 = st + SL.s2; c.m2 = st + SL.m2; c.s3 = st + SL.s3; c.m3 = st + SL.m3;
    c.s5 = st + SL.s5; c.m5 = st + SL.m5; c.sv = st + SL.sv; c.mv = st + SL.mv; c.sg = st + SL.sg; c.mg = st + SL.mg;
    c.scal = st + SL.scal;
    c.G = scr + P.SC.G; c.Minv = scr + P.SC.Minv;
    c.cx = cxp; c.lb = clb; c.ub = cub; c.al = cal; c.be = cbe; c.rhs = rhs; c.cw = cw; c.gc = gc;
    F.mbar_phase = 0;
    F.t_start = globaltimer_ns();
#if OMC_USE_TMA
    mbar_init(mbar, 1);
#endif
  }
  __syncthreads();

  for (;;) {
    // ---------------------------------------------------------------- next node from the queue
    __syncthreads();
    if (tid == 0) {
      const int node = atomicAdd(P.queue, 1);
      F.node = node;
      ish[3] = 0;
      if (node < P.B) {
        const int L = P.node_cut_ptr[node + 1] - P.node_cut_ptr[node];
        F.c.L = L;
        F.c.r = 1 + L * (k + 1);
      }
    }
    __syncthreads();
    const int node = F.node;
    if (node >= P.B) break;
    if (P.prof)
      for (int q = tid; q < OMC_PROF_STRIDE; q += NT) P.prof[(size_t)node * OMC_PROF_STRIDE + q] = 0.0;
    relax_node_setup<NT, KMAX, PM>(P, F);
    for (int it = 1; it <= P.o.max_iter; ++it) {
      if (tid == 0) F.it = it;
      __syncthreads();
#ifdef OMC_INFEASIBILITY_CERTIFICATE
      if (F.c.L > 0 && (it % P.o.check_every == 0 || it == P.o.max_iter || F.force_check)) relax_save_mu<NT, KMAX, PM>(P, F);
#endif
      relax_phase12<NT, KMAX, PM>(P, F);
      for (int b = 0; b < 3; ++b) relax_project_block<NT, KMAX, PM>(P, F, b);
      if (it % P.o.check_every == 0 || it == P.o.max_iter || F.force_check) {
        if (relax_phase4<NT, KMAX, PM>(P, F)) break;
      }
    }
    relax_node_output<NT, KMAX, PM>(P, F);
  }
}


// shared memory bytes the kernel carves up (must mirror the carve-up above)
template <int PM>
inline size_t relax_smem_bytes(int n, int m, int k, int Lcap, int rmax, int xs_cap) {
  Geo g1 = make_geo(n + m);
  Geo gf = smem_geo(n + m, n + k, n);
  size_t d = 0;
  d += (size_t)gf.NP * gf.ld + region1_doubles<PM>(gf, g1.NP);   // buf0, region 1
  const size_t NPI = g1.NP < 16 ? 16 : g1.NP;
  d += 2 * NPI;                          // lam, wgt
  d += 2 * (NPI / 2);                    // jcs, jsn
  d += 32;                               // red
  d += 3 * (size_t)rmax;                 // rhs, cw, gc
  d += 3 * (size_t)Lcap * k + Lcap;      // clb, cub, cal, cbe
  d += (size_t)Lcap;                     // cxp (pointers, 8 bytes)
  d += (size_t)Lcap + (size_t)xs_cap;    // tgs, xs
  d += 1 + 24;                           // mbar, sprof
  d += (sizeof(KFrame) + 7) / 8;         // the shared frame
  size_t bytes = d * 8;
  bytes += sizeof(int) * (3 * (NPI / 2) + 2 * NPI + 8);
  return (bytes + 127) & ~(size_t)127;
}

}  // namespace omc
