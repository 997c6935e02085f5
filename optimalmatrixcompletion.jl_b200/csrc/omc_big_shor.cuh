// omc_big_shor.cuh -- Shor valid-inequality rows (K4) of the node relaxation inside the batched engine:
// OMC.jl:1503-1552 (variables), 1755-1828 (rows), 1838-1846 (objective), for any k.  CPU restatement with the same
// elimination and the same closed-form w-update: oracle/shor_relax.py.
//
//   X = sum_t Xt[t];  W = sum_t Wd[t] + 2 sum_{t1<t2} H[(t1,t2)]  (covered coordinates; an uncovered coordinate keeps W = Wd[0])
//   rows:  5 x 5 moment block per (minor, t)             PSD   (360 356 blocks at config 3)
//          (k+1) x (k+1) block per covered coordinate    PSD   (k > 1)
//          (1/2, W, X) per uncovered coordinate          RSOC
//          Wd >= 0;   a Theta~_jj - sum_i W_ij = 0       zero cone, one row per column j
//   objective  1/2 sum_I (A^2 - 2 A X + W) + cT tr Theta~   (linear: P = 0)
//
// Every row is a selection except the X part of the big block (sum over t: per-coordinate Sherman-Morrison, in k_xt) and the
// zero-cone rows (per-column Sherman-Morrison, k_shor_col).  The 5 x 5 blocks live in HBM in entry-major (SoA) order
// [t][entry 0..14][minor], so that the projection / update kernels (one thread per block) read and write coalesced; the
// per-variable sums over incident blocks go through CSR incidence lists built once per problem (host side, omc_big.cu).
#pragma once
#include "omc_big.cuh"

namespace omcbig {

// packed lower-triangular index of a 5 x 5 block: (r, c), r >= c
__host__ __device__ constexpr int p5(int r, int c) { return r * (r + 1) / 2 + c; }

// Jacobi eigendecomposition of a symmetric matrix of order <= 5 in registers; returns PSD part in S (packed, 15 entries).
__device__ __forceinline__ void psd5(const double (&v)[B5], double (&s)[B5]) {
  double a[5][5], q[5][5];
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) { a[r][c] = v[r >= c ? p5(r, c) : p5(c, r)]; q[r][c] = (r == c) ? 1.0 : 0.0; }
  double scale = 0.0;
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) scale = fmax(scale, fabs(a[r][c]));
  if (scale > 0.0) {
    for (int sweep = 0; sweep < 10; ++sweep) {
      double off = 0.0;
#pragma unroll
      for (int r = 1; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < r; ++c) off = fmax(off, fabs(a[r][c]));
      if (off <= 1e-15 * scale) break;
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int qq = p + 1; qq < 5; ++qq) {
          const double apq = a[p][qq];
          if (fabs(apq) > 1e-300 && fabs(apq) > 1e-18 * scale) {
            const double tau = (a[qq][qq] - a[p][p]) / (2.0 * apq);
            const double t = ((tau >= 0.0) ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            const double c_ = rsqrt(1.0 + t * t), s_ = t * c_;
#pragma unroll
            for (int r = 0; r < 5; ++r) {          // columns p, q
              const double x = a[r][p], y = a[r][qq];
              a[r][p] = c_ * x - s_ * y; a[r][qq] = s_ * x + c_ * y;
              const double u = q[r][p], w = q[r][qq];
              q[r][p] = c_ * u - s_ * w; q[r][qq] = s_ * u + c_ * w;
            }
#pragma unroll
            for (int r = 0; r < 5; ++r) {          // rows p, q
              const double x = a[p][r], y = a[qq][r];
              a[p][r] = c_ * x - s_ * y; a[qq][r] = s_ * x + c_ * y;
            }
          }
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int e = 0; e < 5; ++e) acc = fma(fmax(a[e][e], 0.0) * q[r][e], q[c][e], acc);
      s[p5(r, c)] = acc;
    }
}

__device__ __forceinline__ void rsoc3(const double (&v)[3], double (&p)[3]) {
  const double r2 = 0.70710678118654752440;
  const double u = (v[0] + v[1]) * r2, w = (v[0] - v[1]) * r2, x = v[2];
  const double nr = sqrt(w * w + x * x);
  double ou, sc;
  if (nr <= u) { ou = u; sc = 1.0; }
  else if (nr <= -u) { ou = 0.0; sc = 0.0; }
  else { ou = 0.5 * (u + nr); sc = (nr > 0.0) ? ou / nr : 0.0; }
  const double w2 = sc * w, x2 = sc * x;
  p[0] = (ou + w2) * r2; p[1] = (ou - w2) * r2; p[2] = x2;
}

__device__ __forceinline__ double* shor_ptr(const ShorDev& sh, int slot) { return sh.SS + (size_t)slot * sh.SL.total; }

// ---- S-a: t = v - 2 P(v) for the 5 x 5 blocks (mode 0) or v - P(v) (mode 1, residual check) ------------------------------
__global__ void __launch_bounds__(128) k_shor_proj5(BigArgs a, int mode) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)SL.k * SL.nm) return;
  const long long t = idx / SL.nm, mi = idx - t * SL.nm;
  const double* vb = Q + SL.vB + (size_t)t * B5 * SL.nm + mi;
  double* tb = Q + SL.TB + (size_t)t * B5 * SL.nm + mi;
  double v[B5], s[B5];
#pragma unroll
  for (int e = 0; e < B5; ++e) v[e] = vb[(size_t)e * SL.nm];
  psd5(v, s);
  const double f = mode ? 1.0 : 2.0;
#pragma unroll
  for (int e = 0; e < B5; ++e) tb[(size_t)e * SL.nm] = v[e] - f * s[e];
}

// ---- S-a': (k+1) blocks and RSOC rows per coordinate ----------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_shor_proj9(BigArgs a, int mode) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= SL.C) return;
  const unsigned char fl = sh.flags[c];
  const double f = mode ? 1.0 : 2.0;
  if ((fl & 1) && SL.k > 1) {
    const int nk = SL.k + 1;
    double v[B5], s[B5];
#pragma unroll
    for (int e = 0; e < B5; ++e) v[e] = 0.0;
    for (int r = 0; r < nk; ++r)
      for (int cc = 0; cc <= r; ++cc) v[p5(r, cc)] = Q[SL.v9 + (size_t)c * SL.K9 + p5(r, cc)];
    psd5(v, s);
    for (int r = 0; r < nk; ++r)
      for (int cc = 0; cc <= r; ++cc) Q[SL.T9 + (size_t)c * SL.K9 + p5(r, cc)] = v[p5(r, cc)] - f * s[p5(r, cc)];
  }
  if (fl & 2) {
    double v[3], p[3];
    for (int e = 0; e < 3; ++e) v[e] = Q[SL.vs + (size_t)c * 3 + e];
    rsoc3(v, p);
    for (int e = 0; e < 3; ++e) Q[SL.Ts + (size_t)c * 3 + e] = v[e] - f * p[e];
  }
}

// ---- S-b: adjoint sums per coordinate (one warp per coordinate): GX[t], GW[t], GH[p] (all "/ rho") ------------------------
//   GX[t,c] = -2 sum_inc TB[t][x-slot] - 2 T9[(1+t,0)] cov - Ts[2] soc            (the big block's X part is added in k_xt)
//   GW[t,c] = -  sum_inc TB[t][w-slot] -   T9[(1+t,1+t)] cov - t7 - [t = 0] Ts[1] soc + t6_j
//   GH[p,c] = -2 T9[(1+t2,1+t1)] cov + 2 t6_j
__global__ void __launch_bounds__(256) k_shor_gather_c(BigArgs a, int mode) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const int lane = threadIdx.x & 31;
  const long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= SL.C) return;
  const int j = (int)(c % a.L.m);
  const unsigned char fl = sh.flags[c];
  const int b0 = sh.cptr[c], b1 = sh.cptr[c + 1];
  const int xs[4] = {p5(1, 0), p5(2, 0), p5(3, 0), p5(4, 0)}, ws[4] = {p5(1, 1), p5(2, 2), p5(3, 3), p5(4, 4)};
  const double f = mode ? 1.0 : 2.0;
  const double t6 = Q[SL.v6 + j];
  for (int t = 0; t < SL.k; ++t) {
    double gx = 0.0, gw = 0.0;
    const double* tb = Q + SL.TB + (size_t)t * B5 * SL.nm;
    for (int e = b0 + lane; e < b1; e += 32) {
      const int code = sh.cinc[e], mi = code >> 2, sl = code & 3;
      gx += tb[(size_t)xs[sl] * SL.nm + mi];
      gw += tb[(size_t)ws[sl] * SL.nm + mi];
    }
    gx = warp_sum(gx); gw = warp_sum(gw);
    if (lane == 0) {
      double GX = -2.0 * gx, GW = -gw;
      if ((fl & 1) && SL.k > 1) {
        GX += -2.0 * Q[SL.T9 + (size_t)c * SL.K9 + p5(1 + t, 0)];
        GW += -Q[SL.T9 + (size_t)c * SL.K9 + p5(1 + t, 1 + t)];
      }
      const bool act = (fl & 1) || ((fl & 2) && t == 0);
      if (act) {
        const double v7 = Q[SL.v7 + (size_t)t * SL.C + c];
        GW += -(v7 - f * fmax(v7, 0.0));
      }
      if (fl & 2) {
        GX += -Q[SL.Ts + (size_t)c * 3 + 2];
        if (t == 0) GW += -Q[SL.Ts + (size_t)c * 3 + 1];
      }
      GW += t6;
      Q[SL.GX + (size_t)t * SL.C + c] = GX;
      Q[SL.GW + (size_t)t * SL.C + c] = act ? GW : 0.0;
    }
  }
  if (lane == 0 && SL.npair > 0) {
    int pi = 0;
    for (int t1 = 0; t1 < SL.k; ++t1)
      for (int t2 = t1 + 1; t2 < SL.k; ++t2, ++pi) {
        double GH = 2.0 * t6;
        if (fl & 1) GH += -2.0 * Q[SL.T9 + (size_t)c * SL.K9 + p5(1 + t2, 1 + t1)];
        Q[SL.GH + (size_t)pi * SL.C + c] = (fl & 1) ? GH : 0.0;
      }
  }
}

// ---- S-b': V1, V2, V3 (pure diagonal): w~ = (sig w + rho g) / (sig + 2 rho cnt)   [V3: 4 rho] ------------------------------
// mode 1: writes |rho g| maxima (stationarity residual; q = 0 for these variables) into chk[0] of the node instead
__global__ void __launch_bounds__(256) k_shor_gather_v(BigArgs a, int mode) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const double* S = node_ptr(a, slot);
  const ShorLayout& SL = sh.SL;
  const double rho = S[a.L.scal + S_RHO], sig = a.o.sigma;
  const long long tot = SL.nv1 + SL.nv2 + SL.nm;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double rdl = 0.0;
  if (idx < tot) {
    for (int t = 0; t < SL.k; ++t) {
      const double* tb = Q + SL.TB + (size_t)t * B5 * SL.nm;
      if (idx < SL.nv1) {
        double g = 0.0;
        const int b0 = sh.v1ptr[idx], b1 = sh.v1ptr[idx + 1];
        for (int e = b0; e < b1; ++e) { const int code = sh.v1inc[e]; g += tb[(size_t)((code & 1) ? p5(4, 3) : p5(2, 1)) * SL.nm + (code >> 1)]; }
        g *= -2.0;
        if (mode) rdl = fmax(rdl, fabs(rho * g));
        else Q[SL.V1t + (size_t)t * SL.nv1 + idx] = (sig * Q[SL.V1 + (size_t)t * SL.nv1 + idx] + rho * g) / (sig + 2.0 * rho * (b1 - b0));
      } else if (idx < SL.nv1 + SL.nv2) {
        const long long id = idx - SL.nv1;
        double g = 0.0;
        const int b0 = sh.v2ptr[id], b1 = sh.v2ptr[id + 1];
        for (int e = b0; e < b1; ++e) { const int code = sh.v2inc[e]; g += tb[(size_t)((code & 1) ? p5(4, 2) : p5(3, 1)) * SL.nm + (code >> 1)]; }
        g *= -2.0;
        if (mode) rdl = fmax(rdl, fabs(rho * g));
        else Q[SL.V2t + (size_t)t * SL.nv2 + id] = (sig * Q[SL.V2 + (size_t)t * SL.nv2 + id] + rho * g) / (sig + 2.0 * rho * (b1 - b0));
      } else {
        const long long mi = idx - SL.nv1 - SL.nv2;
        const double g = -2.0 * (tb[(size_t)p5(4, 1) * SL.nm + mi] + tb[(size_t)p5(3, 2) * SL.nm + mi]);
        if (mode) rdl = fmax(rdl, fabs(rho * g));
        else Q[SL.V3t + (size_t)t * SL.nm + mi] = (sig * Q[SL.V3 + (size_t)t * SL.nm + mi] + rho * g) / (sig + 4.0 * rho);
      }
    }
  }
  if (mode) {
    __shared__ double red[32];
    rdl = block_max(rdl, red);
    if (threadIdx.x == 0 && rdl > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(Q + SL.chk + 0), __double_as_longlong(rdl));
  }
}

// ---- S-c: per column j: Wd~, H~, Theta~_jj by Sherman-Morrison over the zero-cone row  a Theta~_jj - sum_i W_ij = 0;
//      then the v-update of that row, and the Theta~_jj entries of the big block (k_xt left them to this kernel) --------------
__global__ void __launch_bounds__(128) k_shor_col(BigArgs a) {
  const ShorDev& sh = a.sh;
  __shared__ double red[32];
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  double* S = node_ptr(a, slot);
  const ShorLayout& SL = sh.SL;
  const Layout& L = a.L;
  const int j = blockIdx.x, n = L.n, m = L.m, N1 = L.N[0], k = SL.k;
  const double rho = S[L.scal + S_RHO], sig = a.o.sigma, al = a.o.alpha, aa = a.a;
  const double e9 = (k > 1) ? 1.0 : 0.0;
  double sWr = 0.0, sWu = 0.0, sHr = 0.0, sHu = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const size_t c = (size_t)i * m + j;
    const unsigned char fl = sh.flags[c];
    const double mk = (double)a.Mk[c];
    for (int t = 0; t < k; ++t) {
      const bool act = (fl & 1) || ((fl & 2) && t == 0);
      if (!act) continue;
      const double dW = sig + rho * (sh.cnt[c] + e9 * (fl & 1) + 1.0 + (((fl & 2) && t == 0) ? 1.0 : 0.0));
      const double rW = sig * Q[SL.Wd + (size_t)t * SL.C + c] - 0.5 * mk + rho * Q[SL.GW + (size_t)t * SL.C + c];
      sWr += rW / dW; sWu += 1.0 / dW;
    }
    if (fl & 1)
      for (int pi = 0; pi < SL.npair; ++pi) {
        const double dH = sig + 2.0 * rho;
        const double rH = sig * Q[SL.H + (size_t)pi * SL.C + c] - mk + rho * Q[SL.GH + (size_t)pi * SL.C + c];
        sHr += rH / dH; sHu += 1.0 / dH;
      }
  }
  sWr = block_sum(sWr, red); sWu = block_sum(sWu, red); sHr = block_sum(sHr, red); sHu = block_sum(sHu, red);
  const double dTd = sig + rho;
  const double v6 = Q[SL.v6 + j];
  const double rTd = sig * S[L.T + (size_t)j * m + j] - a.cT + Q[SL.gTd + j] - rho * aa * v6;    // gTd = -rho (D - 2F) from k_xt
  const double uDr = aa * rTd / dTd - sWr - 2.0 * sHr;
  const double uDu = aa * aa / dTd + sWu + 4.0 * sHu;
  const double coef = rho * uDr / (1.0 + rho * uDu);
  const double Tdt = rTd / dTd - aa * coef / dTd;
  double sumW = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const size_t c = (size_t)i * m + j;
    const unsigned char fl = sh.flags[c];
    const double mk = (double)a.Mk[c];
    for (int t = 0; t < k; ++t) {
      const bool act = (fl & 1) || ((fl & 2) && t == 0);
      double wt = 0.0;
      if (act) {
        const double dW = sig + rho * (sh.cnt[c] + e9 * (fl & 1) + 1.0 + (((fl & 2) && t == 0) ? 1.0 : 0.0));
        const double rW = sig * Q[SL.Wd + (size_t)t * SL.C + c] - 0.5 * mk + rho * Q[SL.GW + (size_t)t * SL.C + c];
        wt = (rW + coef) / dW;
      }
      Q[SL.Wdt + (size_t)t * SL.C + c] = wt;
      sumW += wt;
    }
    for (int pi = 0; pi < SL.npair; ++pi) {
      double ht = 0.0;
      if (fl & 1) {
        const double dH = sig + 2.0 * rho;
        const double rH = sig * Q[SL.H + (size_t)pi * SL.C + c] - mk + rho * Q[SL.GH + (size_t)pi * SL.C + c];
        ht = (rH + 2.0 * coef) / dH;
      }
      Q[SL.Ht + (size_t)pi * SL.C + c] = ht;
      sumW += 2.0 * ht;
    }
  }
  sumW = block_sum(sumW, red);
  if (threadIdx.x == 0) {
    Q[SL.v6 + j] = v6 + al * (aa * Tdt - sumW);                    // zero cone: s = 0
    const size_t ev = L.V[0] + (size_t)(n + j) * N1 + n + j;
    S[ev] = S[ev] + al * (Tdt - Q[SL.Fd + j]);                      // big block, Theta~_jj entry
    const size_t et = L.T + (size_t)j * m + j;
    S[et] = al * Tdt + (1.0 - al) * S[et];
  }
}

// ---- S-d: v-update of the 5 x 5 blocks: v += alpha (z~ - s), s = (v - TB) / 2, z~ gathered from w~ ---------------------------
__global__ void __launch_bounds__(128) k_shor_v5(BigArgs a) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)SL.k * SL.nm) return;
  const long long t = idx / SL.nm, mi = idx - t * SL.nm;
  const int m = a.L.m;
  const int i1 = sh.minors[4 * mi], i2 = sh.minors[4 * mi + 1], j1 = sh.minors[4 * mi + 2], j2 = sh.minors[4 * mi + 3];
  const size_t cc[4] = {(size_t)i1 * m + j1, (size_t)i1 * m + j2, (size_t)i2 * m + j1, (size_t)i2 * m + j2};
  double z[B5];
  z[p5(0, 0)] = 1.0;
  const double* Xtt = Q + SL.Xtt + (size_t)t * SL.C; const double* Wdt = Q + SL.Wdt + (size_t)t * SL.C;
#pragma unroll
  for (int s = 0; s < 4; ++s) { z[p5(1 + s, 0)] = Xtt[cc[s]]; z[p5(1 + s, 1 + s)] = Wdt[cc[s]]; }
  const int* mv = sh.mv + 4 * mi;
  z[p5(2, 1)] = Q[SL.V1t + (size_t)t * SL.nv1 + mv[0]]; z[p5(4, 3)] = Q[SL.V1t + (size_t)t * SL.nv1 + mv[1]];
  z[p5(3, 1)] = Q[SL.V2t + (size_t)t * SL.nv2 + mv[2]]; z[p5(4, 2)] = Q[SL.V2t + (size_t)t * SL.nv2 + mv[3]];
  const double v3 = Q[SL.V3t + (size_t)t * SL.nm + mi];
  z[p5(4, 1)] = v3; z[p5(3, 2)] = v3;
  double* vb = Q + SL.vB + (size_t)t * B5 * SL.nm + mi;
  const double* tb = Q + SL.TB + (size_t)t * B5 * SL.nm + mi;
  const double al = a.o.alpha;
#pragma unroll
  for (int e = 0; e < B5; ++e) {
    const double v = vb[(size_t)e * SL.nm], s = 0.5 * (v - tb[(size_t)e * SL.nm]);
    vb[(size_t)e * SL.nm] = v + al * (z[e] - s);
  }
}

// ---- S-d': v-update of the per-coordinate rows ((k+1) blocks, RSOC, Wd >= 0) and relaxation of the coordinate variables ----
__global__ void __launch_bounds__(256) k_shor_vc(BigArgs a) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= SL.C) return;
  const unsigned char fl = sh.flags[c];
  const double al = a.o.alpha;
  const int k = SL.k;
  double xs = 0.0;
  for (int t = 0; t < k; ++t) xs += Q[SL.Xtt + (size_t)t * SL.C + c];
  if ((fl & 1) && k > 1) {
    for (int r = 0; r <= k; ++r)
      for (int cc = 0; cc <= r; ++cc) {
        double z;
        if (r == 0) z = 1.0;
        else if (cc == 0) z = Q[SL.Xtt + (size_t)(r - 1) * SL.C + c];
        else if (cc == r) z = Q[SL.Wdt + (size_t)(r - 1) * SL.C + c];
        else {                                     // pair (cc-1, r-1), cc-1 < r-1: index in combinations order
          const int t1 = cc - 1, t2 = r - 1;
          const int pi = t1 * k - t1 * (t1 + 1) / 2 + (t2 - t1 - 1);
          z = Q[SL.Ht + (size_t)pi * SL.C + c];
        }
        const size_t e = SL.v9 + (size_t)c * SL.K9 + p5(r, cc);
        const double v = Q[e], s = 0.5 * (v - Q[SL.T9 + (size_t)c * SL.K9 + p5(r, cc)]);
        Q[e] = v + al * (z - s);
      }
  }
  if (fl & 2) {
    const double z3[3] = {0.5, Q[SL.Wdt + c], xs};
    for (int e = 0; e < 3; ++e) {
      const double v = Q[SL.vs + (size_t)c * 3 + e], s = 0.5 * (v - Q[SL.Ts + (size_t)c * 3 + e]);
      Q[SL.vs + (size_t)c * 3 + e] = v + al * (z3[e] - s);
    }
  }
  for (int t = 0; t < k; ++t) {
    const bool act = (fl & 1) || ((fl & 2) && t == 0);
    const size_t e = (size_t)t * SL.C + c;
    if (act) {
      const double v7 = Q[SL.v7 + e];
      Q[SL.v7 + e] = v7 + al * (Q[SL.Wdt + e] - fmax(v7, 0.0));
    }
    Q[SL.Xt + e] = al * Q[SL.Xtt + e] + (1.0 - al) * Q[SL.Xt + e];
    Q[SL.Wd + e] = al * Q[SL.Wdt + e] + (1.0 - al) * Q[SL.Wd + e];
  }
  for (int pi = 0; pi < SL.npair; ++pi) {
    const size_t e = (size_t)pi * SL.C + c;
    Q[SL.H + e] = al * Q[SL.Ht + e] + (1.0 - al) * Q[SL.H + e];
  }
}

// relaxation of V1, V2, V3 (after k_shor_v5 has read the tilde values)
__global__ void __launch_bounds__(256) k_shor_relax_v(BigArgs a) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const double al = a.o.alpha;
  const long long tot = (long long)SL.k * (SL.nv1 + SL.nv2 + SL.nm);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= tot) return;
  size_t w, wt;
  if (idx < (long long)SL.k * SL.nv1) { w = SL.V1 + idx; wt = SL.V1t + idx; }
  else if (idx < (long long)SL.k * (SL.nv1 + SL.nv2)) { const long long q = idx - (long long)SL.k * SL.nv1; w = SL.V2 + q; wt = SL.V2t + q; }
  else { const long long q = idx - (long long)SL.k * (SL.nv1 + SL.nv2); w = SL.V3 + q; wt = SL.V3t + q; }
  Q[w] = al * Q[wt] + (1.0 - al) * Q[w];
}

// ---- residual check of the Shor rows and variables (after k_shor_proj5/9 and the gathers in mode 1: TB, T9, Ts = v - s = mu / rho,
//      GX / GW / GH = adjoint of mu / rho).  chk[0] = rd (max), chk[1] = rp (max), chk[2] = sum mu00 (dual objective), chk[3] = sum_I W ----
__global__ void __launch_bounds__(256) k_shor_check_c(BigArgs a) {
  const ShorDev& sh = a.sh;
  __shared__ double red[32];
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const double* S = node_ptr(a, slot);
  const ShorLayout& SL = sh.SL;
  const Layout& L = a.L;
  const double rho = S[L.scal + S_RHO];
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double rd = 0.0, rp = 0.0, mu00 = 0.0, sW = 0.0;
  const int k = SL.k;
  if (c < SL.C) {
    const unsigned char fl = sh.flags[c];
    const int i = (int)(c / L.m), j = (int)(c % L.m);
    const double mk = (double)a.Mk[c], am = a.AM[c];
    // X_t: q = -A mask; adjoint = GX + (big block X part: -2 (V1 - F) computed by k_check as mu1; stored by it in gX scratch? no:)
    // the big-block part is added here from the node record: mu1_X / rho = V1[i][n+j] - F1; F1 is not available per entry, so
    // k_check (which has the panels) leaves  -2 rho (V1 - F1)[i, n+j]  in S[L.Yt]-free scratch: we use Q[SL.Xtt] slot 0 as scratch.
    const double big = Q[SL.Xtt + c];
    double W = 0.0, xs = 0.0;
    for (int t = 0; t < k; ++t) {
      const size_t e = (size_t)t * SL.C + c;
      rd = fmax(rd, fabs(-am - (big + rho * Q[SL.GX + e])));
      const bool act = (fl & 1) || ((fl & 2) && t == 0);
      if (act) {
        rd = fmax(rd, fabs(0.5 * mk - rho * Q[SL.GW + e]));
        const double wd = Q[SL.Wd + e], v7 = Q[SL.v7 + e];
        rp = fmax(rp, fabs(wd - fmax(v7, 0.0)));
        W += wd;
      }
      xs += Q[SL.Xt + e];
    }
    for (int pi = 0; pi < SL.npair; ++pi)
      if (fl & 1) { rd = fmax(rd, fabs(mk - rho * Q[SL.GH + (size_t)pi * SL.C + c])); W += 2.0 * Q[SL.H + (size_t)pi * SL.C + c]; }
    sW = mk * W;
    if ((fl & 1) && k > 1) {
      for (int r = 0; r <= k; ++r)
        for (int cc = 0; cc <= r; ++cc) {
          double z;
          if (r == 0) z = 1.0;
          else if (cc == 0) z = Q[SL.Xt + (size_t)(r - 1) * SL.C + c];
          else if (cc == r) z = Q[SL.Wd + (size_t)(r - 1) * SL.C + c];
          else { const int t1 = cc - 1, t2 = r - 1; z = Q[SL.H + (size_t)(t1 * k - t1 * (t1 + 1) / 2 + (t2 - t1 - 1)) * SL.C + c]; }
          const double v = Q[SL.v9 + (size_t)c * SL.K9 + p5(r, cc)], mu = Q[SL.T9 + (size_t)c * SL.K9 + p5(r, cc)];   // mu / rho = v - s
          rp = fmax(rp, fabs(z - (v - mu)));
          if (r == 0) mu00 += rho * mu;
        }
    }
    if (fl & 2) {
      const double z3[3] = {0.5, Q[SL.Wd + c], xs};
      for (int e = 0; e < 3; ++e) {
        const double v = Q[SL.vs + (size_t)c * 3 + e], mu = Q[SL.Ts + (size_t)c * 3 + e];
        rp = fmax(rp, fabs(z3[e] - (v - mu)));
        if (e == 0) mu00 += 0.5 * rho * mu;
      }
    }
    (void)i; (void)j;
  }
  rd = block_max(rd, red); rp = block_max(rp, red); mu00 = block_sum(mu00, red); sW = block_sum(sW, red);
  if (threadIdx.x == 0) {
    if (rd > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(Q + SL.chk + 0), __double_as_longlong(rd));
    if (rp > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(Q + SL.chk + 1), __double_as_longlong(rp));
    atomicAdd(Q + SL.chk + 2, mu00);
    atomicAdd(Q + SL.chk + 3, sW);
  }
}

// 5 x 5 blocks: primal residual |M5(w) - s| and the constant-entry multipliers (dual objective)
__global__ void __launch_bounds__(128) k_shor_check_5(BigArgs a) {
  const ShorDev& sh = a.sh;
  __shared__ double red[32];
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const double* S = node_ptr(a, slot);
  const ShorLayout& SL = sh.SL;
  const double rho = S[a.L.scal + S_RHO];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double rp = 0.0, mu00 = 0.0;
  if (idx < (long long)SL.k * SL.nm) {
    const long long t = idx / SL.nm, mi = idx - t * SL.nm;
    const int m = a.L.m;
    const int i1 = sh.minors[4 * mi], i2 = sh.minors[4 * mi + 1], j1 = sh.minors[4 * mi + 2], j2 = sh.minors[4 * mi + 3];
    const size_t cc[4] = {(size_t)i1 * m + j1, (size_t)i1 * m + j2, (size_t)i2 * m + j1, (size_t)i2 * m + j2};
    double z[B5];
    z[p5(0, 0)] = 1.0;
    for (int s = 0; s < 4; ++s) { z[p5(1 + s, 0)] = Q[SL.Xt + (size_t)t * SL.C + cc[s]]; z[p5(1 + s, 1 + s)] = Q[SL.Wd + (size_t)t * SL.C + cc[s]]; }
    const int* mv = sh.mv + 4 * mi;
    z[p5(2, 1)] = Q[SL.V1 + (size_t)t * SL.nv1 + mv[0]]; z[p5(4, 3)] = Q[SL.V1 + (size_t)t * SL.nv1 + mv[1]];
    z[p5(3, 1)] = Q[SL.V2 + (size_t)t * SL.nv2 + mv[2]]; z[p5(4, 2)] = Q[SL.V2 + (size_t)t * SL.nv2 + mv[3]];
    const double v3 = Q[SL.V3 + (size_t)t * SL.nm + mi];
    z[p5(4, 1)] = v3; z[p5(3, 2)] = v3;
    const double* vb = Q + SL.vB + (size_t)t * B5 * SL.nm + mi;
    const double* tb = Q + SL.TB + (size_t)t * B5 * SL.nm + mi;
#pragma unroll
    for (int e = 0; e < B5; ++e) {
      const double v = vb[(size_t)e * SL.nm], mu = tb[(size_t)e * SL.nm];      // mu / rho = v - s
      rp = fmax(rp, fabs(z[e] - (v - mu)));
      if (e == 0) mu00 += rho * mu;
    }
  }
  rp = block_max(rp, red); mu00 = block_sum(mu00, red);
  if (threadIdx.x == 0) {
    if (rp > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(Q + SL.chk + 1), __double_as_longlong(rp));
    atomicAdd(Q + SL.chk + 2, mu00);
  }
}

// zero-cone rows and Theta~_jj stationarity (one thread per column)
__global__ void __launch_bounds__(128) k_shor_check_col(BigArgs a) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  double* Q = shor_ptr(sh, slot);
  const double* S = node_ptr(a, slot);
  const ShorLayout& SL = sh.SL;
  const Layout& L = a.L;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L.m) return;
  const double rho = S[L.scal + S_RHO];
  double sumW = 0.0;
  for (int i = 0; i < L.n; ++i) {
    const size_t c = (size_t)i * L.m + j;
    const unsigned char fl = sh.flags[c];
    for (int t = 0; t < SL.k; ++t) if ((fl & 1) || ((fl & 2) && t == 0)) sumW += Q[SL.Wd + (size_t)t * SL.C + c];
    if (fl & 1) for (int pi = 0; pi < SL.npair; ++pi) sumW += 2.0 * Q[SL.H + (size_t)pi * SL.C + c];
  }
  const double rp = fabs(a.a * S[L.T + (size_t)j * L.m + j] - sumW);
  // stationarity of Theta~_jj: cT - A'mu = cT + mu1_jj + rho a v6_j;  mu1_jj = -gTd / 1 (k_check stored -(V1 - F) rho in gTd)
  const double rd = fabs(a.cT - Q[SL.gTd + j] + rho * a.a * Q[SL.v6 + j]);
  if (rp > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(Q + SL.chk + 1), __double_as_longlong(rp));
  if (rd > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(Q + SL.chk + 0), __double_as_longlong(rd));
}


// ---- node setup: constant entries of the v-form rows ((0,0) = 1 of every moment block, 1/2 of the RSOC rows) ---------------
__global__ void __launch_bounds__(256) k_shor_init(BigArgs a) {
  const ShorDev& sh = a.sh;
  const int slot = blockIdx.y;
  double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (long long)SL.k * SL.nm) {
    const long long t = idx / SL.nm, mi = idx - t * SL.nm;
    Q[SL.vB + (size_t)t * B5 * SL.nm + mi] = 1.0;          // entry (0,0)
  }
  if (idx < SL.C) {
    const unsigned char fl = sh.flags[idx];
    if ((fl & 1) && SL.k > 1) Q[SL.v9 + (size_t)idx * SL.K9] = 1.0;
    if (fl & 2) Q[SL.vs + (size_t)idx * 3] = 0.5;
  }
}

__global__ void k_shor_zero_chk(BigArgs a) {
  const ShorDev& sh = a.sh;
  double* Q = shor_ptr(sh, a.active[blockIdx.x]);
  if (threadIdx.x < 8) Q[sh.SL.chk + threadIdx.x] = 0.0;
}

// ---- rho change: mu stays, v <- s + cf (v - s) on every Shor row ---------------------------------------------------------------
__global__ void __launch_bounds__(128) k_shor_rescale5(BigArgs a) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  if (!node_int(a, slot)[I_ADAPTED]) return;
  double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const double cf = node_ptr(a, slot)[a.L.scal + S_CFAC];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)SL.k * SL.nm) return;
  const long long t = idx / SL.nm, mi = idx - t * SL.nm;
  double* vb = Q + SL.vB + (size_t)t * B5 * SL.nm + mi;
  double v[B5], s[B5];
#pragma unroll
  for (int e = 0; e < B5; ++e) v[e] = vb[(size_t)e * SL.nm];
  psd5(v, s);
#pragma unroll
  for (int e = 0; e < B5; ++e) vb[(size_t)e * SL.nm] = s[e] + cf * (v[e] - s[e]);
}
__global__ void __launch_bounds__(256) k_shor_rescale_c(BigArgs a) {
  const ShorDev& sh = a.sh;
  const int slot = a.active[blockIdx.y];
  if (!node_int(a, slot)[I_ADAPTED]) return;
  double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const double cf = node_ptr(a, slot)[a.L.scal + S_CFAC];
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < a.L.m) Q[SL.v6 + c] *= cf;
  if (c >= SL.C) return;
  const unsigned char fl = sh.flags[c];
  if ((fl & 1) && SL.k > 1) {
    const int nk = SL.k + 1;
    double v[B5], s[B5];
#pragma unroll
    for (int e = 0; e < B5; ++e) v[e] = 0.0;
    for (int r = 0; r < nk; ++r)
      for (int cc = 0; cc <= r; ++cc) v[p5(r, cc)] = Q[SL.v9 + (size_t)c * SL.K9 + p5(r, cc)];
    psd5(v, s);
    for (int r = 0; r < nk; ++r)
      for (int cc = 0; cc <= r; ++cc) Q[SL.v9 + (size_t)c * SL.K9 + p5(r, cc)] = s[p5(r, cc)] + cf * (v[p5(r, cc)] - s[p5(r, cc)]);
  }
  if (fl & 2) {
    double v[3], p[3];
    for (int e = 0; e < 3; ++e) v[e] = Q[SL.vs + (size_t)c * 3 + e];
    rsoc3(v, p);
    for (int e = 0; e < 3; ++e) Q[SL.vs + (size_t)c * 3 + e] = p[e] + cf * (v[e] - p[e]);
  }
  for (int t = 0; t < SL.k; ++t) {
    const double v7 = Q[SL.v7 + (size_t)t * SL.C + c], s7 = fmax(v7, 0.0);
    Q[SL.v7 + (size_t)t * SL.C + c] = s7 + cf * (v7 - s7);
  }
}

// results: W (column-major n x m) and Xt (k slices, column-major) of every node
__global__ void __launch_bounds__(256) k_shor_extract(BigArgs a, int B, double* outW, double* outXt) {
  const ShorDev& sh = a.sh;
  const int slot = blockIdx.y;
  if (slot >= B) return;
  const double* Q = shor_ptr(sh, slot);
  const ShorLayout& SL = sh.SL;
  const int n = a.L.n, m = a.L.m;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= SL.C) return;
  const int j = (int)(e / n), i = (int)(e - (long long)j * n);
  const size_t c = (size_t)i * m + j;
  double W = 0.0;
  for (int t = 0; t < SL.k; ++t) {
    W += Q[SL.Wd + (size_t)t * SL.C + c];
    if (outXt) outXt[((size_t)slot * SL.k + t) * SL.C + e] = Q[SL.Xt + (size_t)t * SL.C + c];
  }
  for (int pi = 0; pi < SL.npair; ++pi) W += 2.0 * Q[SL.H + (size_t)pi * SL.C + c];
  if (outW) outW[(size_t)slot * SL.C + e] = W;
}

}  // namespace omcbig
