// omc_altmin.cuh -- alternating minimisation (K7 / K8), replacing the body of alternating_minimization
// (/root/reference/src/OptimalMatrixCompletion.jl:1979-2279).  The whole heuristic -- V-steps, U-steps and the
// stopping rule -- runs inside one kernel launch with no host round trip:
//   V-step  (OMC.jl:2192-2209): per column j the exact k x k normal equations
//           (U_Ij' U_Ij + U'U / gamma) v_j = U_Ij' A_Ij,j  over the column-CSC of the mask, Cholesky in registers;
//   U-step  (OMC.jl:2212-2229): min_U f(U, V) subject to the box with symmetry-breaking zeros, ||U_j|| <= 1,
//           ||U_a +- U_b|| <= sqrt 2 and the cut half-spaces (OMC.jl:2024-2091, 2164-2171), solved by ADMM in
//           OSQP form.  The objective is row-separable (H_i = V_Ii V_Ii' + VV'/gamma over the row-CSR), the
//           constraint rows make A'A = 2k I plus a rank L k term, so the linear solve is a batched k x k inverse
//           per row plus a small Woodbury correction for the cut rows.
// One CTA: these systems are latency-bound (n k <= a few thousand unknowns); batching over nodes is the
// "next" row in DESIGN.md.
#pragma once
#include "omc_device.cuh"
#include "omc_relax.cuh"

namespace omc {

constexpr int ALT_MAXK = 8;

struct AltminArgs {
  int n, m, k, L, cut_type, fix3;
  double gamma, eps, inner_eps, sigma, alpha, time_limit_s;
  int max_iters, inner_max;
  const double* A;  // n*m col-major
  const int *rowptr, *colidx, *colptr, *rowidx;
  const double* pool_x;
  const double* pool_vhat;
  const int* cut_ids;        // [L]
  const uint8_t* cut_dirs;   // [L*k]
  double* U;                 // n*k col-major: in = U_initial, out = U
  double* V;                 // k*m col-major (k x m): out
  double* ws;                // workspace (doubles), layout below
  double* objectives;        // [max_iters]
  int* out_int;              // [0] converged, [1] n_iters, [2] total inner iterations
  // batch (one CTA per problem instance: same A / mask / gamma, own U_initial and cut list): block b uses
  // U + b n k, V + b k m, ws + b ws_stride, objectives + b max_iters, out_int + 4 b and cuts cut_ptr[b] .. cut_ptr[b+1]
  const int* cut_ptr;        // [B+1] or nullptr (single instance: the fields above as they are)
  size_t ws_stride;
};

// workspace layout (doubles)
struct AltminWs {
  size_t Ut, H, g, K, zb, yb, zc, yc, zp, yp, zm, ym, zv, yv, lb, ub, M, T, total;
};
__host__ __device__ inline AltminWs make_altmin_ws(int n, int m, int k, int L) {
  AltminWs w;
  size_t o = 0;
  const size_t nk = (size_t)n * k, npair = (size_t)k * (k - 1) / 2, Lk = (size_t)L * k;
  w.Ut = o; o += nk;
  w.H = o; o += nk * k;
  w.g = o; o += nk;
  w.K = o; o += nk * k;
  w.zb = o; o += nk; w.yb = o; o += nk;
  w.zc = o; o += nk; w.yc = o; o += nk;
  w.zp = o; o += npair * n; w.yp = o; o += npair * n;
  w.zm = o; o += npair * n; w.ym = o; o += npair * n;
  w.zv = o; o += Lk; w.yv = o; o += Lk;
  w.lb = o; o += Lk; w.ub = o; o += Lk;
  w.M = o; o += Lk * Lk;
  w.T = o; o += 2 * nk + 2 * Lk;  // scratch: rhs [nk], R u~ [Lk], cw [Lk], K R' cw [nk]
  w.total = o + 8;
  return w;
}

// in-thread Cholesky solve of the SPD k x k system H v = g (H row-major, destroyed); returns false if singular
__device__ __forceinline__ bool chol_solve_small(double* H, double* g, int k) {
  for (int j = 0; j < k; ++j) {
    double d = H[j * k + j];
    for (int p = 0; p < j; ++p) d -= H[j * k + p] * H[j * k + p];
    if (!(d > 1e-300)) return false;
    d = sqrt(d);
    H[j * k + j] = d;
    for (int i = j + 1; i < k; ++i) {
      double s = H[i * k + j];
      for (int p = 0; p < j; ++p) s -= H[i * k + p] * H[j * k + p];
      H[i * k + j] = s / d;
    }
  }
  for (int i = 0; i < k; ++i) {
    double s = g[i];
    for (int p = 0; p < i; ++p) s -= H[i * k + p] * g[p];
    g[i] = s / H[i * k + i];
  }
  for (int i = k - 1; i >= 0; --i) {
    double s = g[i];
    for (int p = i + 1; p < k; ++p) s -= H[p * k + i] * g[p];
    g[i] = s / H[i * k + i];
  }
  return true;
}

// in-thread inverse of the SPD k x k matrix (Gauss-Jordan, no pivoting), in place
__device__ __forceinline__ void inv_small(double* M, int k) {
  for (int p = 0; p < k; ++p) {
    const double inv = 1.0 / M[p * k + p];
    for (int j = 0; j < k; ++j)
      if (j != p) M[p * k + j] *= inv;
    for (int i = 0; i < k; ++i) {
      if (i == p) continue;
      const double f = M[i * k + p];
      for (int j = 0; j < k; ++j)
        if (j != p) M[i * k + j] -= f * M[p * k + j];
      M[i * k + p] = -f * inv;
    }
    M[p * k + p] = inv;
  }
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) altmin_kernel(const AltminArgs P0) {
  AltminArgs P = P0;
  if (P0.cut_ptr != nullptr) {
    const size_t b = blockIdx.x;
    const int c0 = P0.cut_ptr[b];
    P.L = P0.cut_ptr[b + 1] - c0;
    P.cut_ids = P0.cut_ids + c0;
    P.cut_dirs = P0.cut_dirs + (size_t)c0 * P0.k;
    P.U = P0.U + b * (size_t)P0.n * P0.k;
    P.V = P0.V + b * (size_t)P0.k * P0.m;
    P.ws = P0.ws + b * P0.ws_stride;
    P.objectives = P0.objectives + b * (size_t)P0.max_iters;
    P.out_int = P0.out_int + 4 * b;
  }
  __shared__ double red[32];
  __shared__ double kk1[ALT_MAXK * ALT_MAXK];  // U'U/gamma or VV'/gamma
  __shared__ double kk2[ALT_MAXK * ALT_MAXK];
  __shared__ double nrm[ALT_MAXK * ALT_MAXK];  // norms for the ball projections
  __shared__ int flag[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const int n = P.n, m = P.m, k = P.k, L = P.L;
  const int nk = n * k, npair = k * (k - 1) / 2, Lk = L * k;
  const AltminWs W = make_altmin_ws(n, m, k, L);
  double* Ut = P.ws + W.Ut; double* H = P.ws + W.H; double* g = P.ws + W.g; double* K = P.ws + W.K;
  double* zb = P.ws + W.zb; double* yb = P.ws + W.yb; double* zc = P.ws + W.zc; double* yc = P.ws + W.yc;
  double* zp = P.ws + W.zp; double* yp = P.ws + W.yp; double* zm = P.ws + W.zm; double* ym = P.ws + W.ym;
  double* zv = P.ws + W.zv; double* yv = P.ws + W.yv; double* clb = P.ws + W.lb; double* cub = P.ws + W.ub;
  double* Mw = P.ws + W.M; double* T1 = P.ws + W.T; double* T2 = T1 + nk; double* T3 = T2 + Lk; double* T4 = T3 + Lk;
  double* U = P.U; double* V = P.V;
  const double r2 = sqrt(2.0);
  const unsigned long long t_start = globaltimer_ns();

  // cut rows (original units; alt-min uses only lb <= x'U_j <= ub, OMC.jl:2049-2091)
  for (int e = tid; e < Lk; e += NT) {
    const int l = e / k, j = e - l * k;
    double lb, ub, al, be;
    cut_coeffs(P.cut_type, P.cut_dirs[e], P.pool_vhat[(size_t)P.cut_ids[l] * k + j], P.fix3, lb, ub, al, be);
    clb[e] = lb; cub[e] = ub;
  }
  for (int e = tid; e < nk; e += NT) { zb[e] = yb[e] = zc[e] = yc[e] = 0.0; }
  for (int e = tid; e < npair * n; e += NT) { zp[e] = yp[e] = zm[e] = ym[e] = 0.0; }
  for (int e = tid; e < Lk; e += NT) { zv[e] = yv[e] = 0.0; }
  double rho = 1.0;
  double objective_current = 1e10;
  int counter = 0, converged = 0, inner_total = 0;
  __syncthreads();

  while (counter < P.max_iters) {
    if (P.time_limit_s > 0.0) {
      if (tid == 0) flag[0] = ((double)(globaltimer_ns() - t_start) * 1e-9 >= P.time_limit_s) ? 1 : 0;
      __syncthreads();
      if (flag[0]) break;
    }
    ++counter;
    // ------------------------------------------------------------------ V-step
    for (int e = tid; e < k * k; e += NT) kk1[e] = 0.0;
    __syncthreads();
    for (int e = warp; e < k * k; e += NW) {  // U'U / gamma
      const int a = e / k, b = e - a * k;
      double s = 0.0;
      for (int i = lane; i < n; i += 32) s += U[i + (size_t)n * a] * U[i + (size_t)n * b];
      s = warp_sum(s);
      if (lane == 0) kk1[e] = s / P.gamma;
    }
    __syncthreads();
    for (int j = tid; j < m; j += NT) {
      double Hj[ALT_MAXK * ALT_MAXK], gj[ALT_MAXK];
      for (int e = 0; e < k * k; ++e) Hj[e] = kk1[e];
      for (int a = 0; a < k; ++a) gj[a] = 0.0;
      for (int q = P.colptr[j]; q < P.colptr[j + 1]; ++q) {
        const int i = P.rowidx[q];
        const double aij = P.A[i + (size_t)n * j];
        double u[ALT_MAXK];
        for (int a = 0; a < k; ++a) u[a] = U[i + (size_t)n * a];
        for (int a = 0; a < k; ++a) {
          gj[a] += u[a] * aij;
          for (int b = 0; b < k; ++b) Hj[a * k + b] += u[a] * u[b];
        }
      }
      if (!chol_solve_small(Hj, gj, k))
        for (int a = 0; a < k; ++a) gj[a] = 0.0;
      for (int a = 0; a < k; ++a) V[a + (size_t)k * j] = gj[a];
    }
    __syncthreads();
    // ------------------------------------------------------------------ U-step setup: H_i, g_i
    for (int e = warp; e < k * k; e += NW) {  // VV' / gamma
      const int a = e / k, b = e - a * k;
      double s = 0.0;
      for (int j = lane; j < m; j += 32) s += V[a + (size_t)k * j] * V[b + (size_t)k * j];
      s = warp_sum(s);
      if (lane == 0) kk2[e] = s / P.gamma;
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      double Hi[ALT_MAXK * ALT_MAXK], gi[ALT_MAXK];
      for (int e = 0; e < k * k; ++e) Hi[e] = kk2[e];
      for (int a = 0; a < k; ++a) gi[a] = 0.0;
      for (int q = P.rowptr[i]; q < P.rowptr[i + 1]; ++q) {
        const int j = P.colidx[q];
        const double aij = P.A[i + (size_t)n * j];
        double v[ALT_MAXK];
        for (int a = 0; a < k; ++a) v[a] = V[a + (size_t)k * j];
        for (int a = 0; a < k; ++a) {
          gi[a] += v[a] * aij;
          for (int b = 0; b < k; ++b) Hi[a * k + b] += v[a] * v[b];
        }
      }
      for (int e = 0; e < k * k; ++e) H[(size_t)i * k * k + e] = Hi[e];
      for (int a = 0; a < k; ++a) g[(size_t)i * k + a] = gi[a];
    }
    __syncthreads();
    // ------------------------------------------------------------------ U-step: ADMM
    bool refactor = true;
    int it = 0;
    for (it = 1; it <= P.inner_max; ++it) {
      if (refactor) {
        const double c = P.sigma + 2.0 * k * rho;
        for (int i = tid; i < n; i += NT) {
          double Ki[ALT_MAXK * ALT_MAXK];
          for (int e = 0; e < k * k; ++e) Ki[e] = H[(size_t)i * k * k + e] + (((e / k) == (e % k)) ? c : 0.0);
          inv_small(Ki, k);
          for (int e = 0; e < k * k; ++e) K[(size_t)i * k * k + e] = Ki[e];
        }
        __syncthreads();
        if (L > 0) {  // M[(l,j),(l',j')] = sum_i x_l[i] x_l'[i] K_i[j,j'] + delta / rho
          for (int e = warp; e < Lk * Lk; e += NW) {
            const int r = e / Lk, c2 = e - r * Lk;
            const int l1 = r / k, j1 = r - l1 * k, l2 = c2 / k, j2 = c2 - l2 * k;
            const double* x1 = P.pool_x + (size_t)P.cut_ids[l1] * n;
            const double* x2 = P.pool_x + (size_t)P.cut_ids[l2] * n;
            double s = 0.0;
            for (int i = lane; i < n; i += 32) s += x1[i] * x2[i] * K[(size_t)i * k * k + j1 * k + j2];
            s = warp_sum(s);
            if (lane == 0) Mw[e] = s + ((r == c2) ? 1.0 / rho : 0.0);
          }
          __syncthreads();
          spd_invert(Mw, Lk);
        }
        refactor = false;
      }
      // rhs = sigma u + g + A'(rho z - y)  -> T1 (n x k, row-major i*k + j), then u~ = K rhs
      for (int e = tid; e < nk; e += NT) {
        const int i = e / k, j = e - i * k;
        const size_t c = (size_t)i + (size_t)n * j;  // col-major index of (i, j)
        double r = P.sigma * U[c] + g[e] + (rho * zb[c] - yb[c]) + (rho * zc[c] - yc[c]);
        int p = 0;
        for (int a = 0; a < k - 1; ++a)
          for (int b = a + 1; b < k; ++b, ++p) {
            if (a != j && b != j) continue;
            const double tp = rho * zp[(size_t)p * n + i] - yp[(size_t)p * n + i];
            const double tm = rho * zm[(size_t)p * n + i] - ym[(size_t)p * n + i];
            r += (a == j) ? (tp + tm) : (tp - tm);
          }
        for (int l = 0; l < L; ++l) r += P.pool_x[(size_t)P.cut_ids[l] * n + i] * (rho * zv[l * k + j] - yv[l * k + j]);
        T1[e] = r;
      }
      __syncthreads();
      for (int e = tid; e < nk; e += NT) {
        const int i = e / k, j = e - i * k;
        double s = 0.0;
        for (int q = 0; q < k; ++q) s += K[(size_t)i * k * k + j * k + q] * T1[(size_t)i * k + q];
        Ut[(size_t)i + (size_t)n * j] = s;
      }
      __syncthreads();
      if (L > 0) {  // Woodbury: u~ -= K R' Minv (R u~)
        for (int e = warp; e < Lk; e += NW) {
          const int l = e / k, j = e - l * k;
          const double* x = P.pool_x + (size_t)P.cut_ids[l] * n;
          double s = 0.0;
          for (int i = lane; i < n; i += 32) s += x[i] * Ut[(size_t)i + (size_t)n * j];
          s = warp_sum(s);
          if (lane == 0) T2[e] = s;
        }
        __syncthreads();
        for (int e = tid; e < Lk; e += NT) {
          double s = 0.0;
          for (int q = 0; q < Lk; ++q) s += Mw[(size_t)e * Lk + q] * T2[q];
          T3[e] = s;  // cw
        }
        __syncthreads();
        for (int e = tid; e < nk; e += NT) {
          const int i = e / k, j = e - i * k;
          double s = 0.0;
          for (int q = 0; q < k; ++q) {
            double rq = 0.0;  // (R' cw)[i, q] = sum_l x_l[i] cw[l, q]
            for (int l = 0; l < L; ++l) rq += P.pool_x[(size_t)P.cut_ids[l] * n + i] * T3[l * k + q];
            s += K[(size_t)i * k * k + j * k + q] * rq;
          }
          T4[e] = s;
        }
        __syncthreads();
        for (int e = tid; e < nk; e += NT) {
          const int i = e / k, j = e - i * k;
          Ut[(size_t)i + (size_t)n * j] -= T4[e];
        }
        __syncthreads();
      }
      // u <- alpha u~ + (1 - alpha) u ; box rows
      const double al = P.alpha;
      for (int e = tid; e < nk; e += NT) {
        const int i = e % n, j = e / n;
        const double ut = Ut[e];
        U[e] = al * ut + (1.0 - al) * U[e];
        const double v = al * ut + (1.0 - al) * zb[e] + yb[e] / rho;
        const double lo = (i >= n - k + j) ? 0.0 : -1.0;  // OMC.jl:1989-1996
        const double zn = fmin(fmax(v, lo), 1.0);
        yb[e] = rho * (v - zn);  // = y + rho (alpha z~ + (1-alpha) z - zn)
        zb[e] = zn;
      }
      // ball rows: columns (radius 1), pairs (radius sqrt 2).  v is staged in place of z, then scaled.
      for (int e = tid; e < nk; e += NT) zc[e] = al * Ut[e] + (1.0 - al) * zc[e] + yc[e] / rho;
      for (int e = tid; e < npair * n; e += NT) {
        const int p = e / n, i = e - p * n;
        int a = 0, b = 1, q = 0;
        for (int a2 = 0; a2 < k - 1; ++a2)
          for (int b2 = a2 + 1; b2 < k; ++b2, ++q)
            if (q == p) { a = a2; b = b2; }
        const double ua = Ut[i + (size_t)n * a], ub = Ut[i + (size_t)n * b];
        zp[e] = al * (ua + ub) + (1.0 - al) * zp[e] + yp[e] / rho;
        zm[e] = al * (ua - ub) + (1.0 - al) * zm[e] + ym[e] / rho;
      }
      __syncthreads();
      for (int e = warp; e < k + 2 * npair; e += NW) {
        const double* v = (e < k) ? (zc + (size_t)e * n) : ((e < k + npair) ? (zp + (size_t)(e - k) * n) : (zm + (size_t)(e - k - npair) * n));
        double s = 0.0;
        for (int i = lane; i < n; i += 32) s += v[i] * v[i];
        s = warp_sum(s);
        if (lane == 0) nrm[e] = sqrt(s);
      }
      __syncthreads();
      for (int e = tid; e < nk; e += NT) {
        const int j = e / n;
        const double v = zc[e], sc = (nrm[j] <= 1.0) ? 1.0 : 1.0 / nrm[j];
        const double zn = v * sc;
        yc[e] = rho * (v - zn);
        zc[e] = zn;
      }
      for (int e = tid; e < npair * n; e += NT) {
        const int p = e / n;
        {
          const double v = zp[e], nr = nrm[k + p], sc = (nr <= r2) ? 1.0 : r2 / nr;
          yp[e] = rho * (v - v * sc);
          zp[e] = v * sc;
        }
        {
          const double v = zm[e], nr = nrm[k + npair + p], sc = (nr <= r2) ? 1.0 : r2 / nr;
          ym[e] = rho * (v - v * sc);
          zm[e] = v * sc;
        }
      }
      // cut rows
      for (int e = warp; e < Lk; e += NW) {
        const int l = e / k, j = e - l * k;
        const double* x = P.pool_x + (size_t)P.cut_ids[l] * n;
        double s = 0.0;
        for (int i = lane; i < n; i += 32) s += x[i] * Ut[(size_t)i + (size_t)n * j];
        s = warp_sum(s);
        if (lane == 0) {
          const double v = al * s + (1.0 - al) * zv[e] + yv[e] / rho;
          const double zn = fmin(fmax(v, clb[e]), cub[e]);
          yv[e] = rho * (v - zn);
          zv[e] = zn;
        }
      }
      __syncthreads();
      // residual check
      if (it % 10 == 0 || it == P.inner_max) {
        double rp = 0.0, rd = 0.0, npr = 1.0, ndr = 1.0;
        for (int e = tid; e < nk; e += NT) {
          const int i = e % n, j = e / n;
          const double u = U[e];
          rp = fmax(rp, fmax(fabs(u - zb[e]), fabs(u - zc[e])));
          npr = fmax(npr, fabs(u));
          double hu = 0.0;
          for (int q = 0; q < k; ++q) hu += H[(size_t)i * k * k + j * k + q] * U[i + (size_t)n * q];
          double gr = hu - g[(size_t)i * k + j] + yb[e] + yc[e];
          int p = 0;
          for (int a = 0; a < k - 1; ++a)
            for (int b = a + 1; b < k; ++b, ++p) {
              if (a == j) gr += yp[(size_t)p * n + i] + ym[(size_t)p * n + i];
              else if (b == j) gr += yp[(size_t)p * n + i] - ym[(size_t)p * n + i];
            }
          for (int l = 0; l < L; ++l) gr += P.pool_x[(size_t)P.cut_ids[l] * n + i] * yv[l * k + j];
          rd = fmax(rd, fabs(gr));
          ndr = fmax(ndr, fmax(fabs(hu), fabs(g[(size_t)i * k + j])));
        }
        for (int e = tid; e < npair * n; e += NT) {
          const int p = e / n, i = e - p * n;
          int a = 0, b = 1, q = 0;
          for (int a2 = 0; a2 < k - 1; ++a2)
            for (int b2 = a2 + 1; b2 < k; ++b2, ++q)
              if (q == p) { a = a2; b = b2; }
          const double ua = U[i + (size_t)n * a], ub = U[i + (size_t)n * b];
          rp = fmax(rp, fmax(fabs(ua + ub - zp[e]), fabs(ua - ub - zm[e])));
        }
        for (int e = warp; e < Lk; e += NW) {
          const int l = e / k, j = e - l * k;
          const double* x = P.pool_x + (size_t)P.cut_ids[l] * n;
          double s = 0.0;
          for (int i = lane; i < n; i += 32) s += x[i] * U[(size_t)i + (size_t)n * j];
          s = warp_sum(s);
          rp = fmax(rp, fabs(s - zv[e]));
        }
        rp = block_max(rp, red);
        rd = block_max(rd, red);
        npr = block_max(npr, red);
        ndr = block_max(ndr, red);
        if (rp <= P.inner_eps * npr && rd <= P.inner_eps * ndr) break;
        if (it % 50 == 0) {
          const double ratio = sqrt((rp / npr) / fmax(rd / ndr, 1e-30));
          if (ratio > 5.0 || ratio < 0.2) {
            rho = fmin(fmax(rho * ratio, 1e-6), 1e6);
            refactor = true;
          }
        }
      }
    }
    inner_total += (it > P.inner_max) ? P.inner_max : it;
    // ------------------------------------------------------------------ objective f(U, V) and the stopping rule
    double sse = 0.0;
    for (int i = warp; i < n; i += NW) {
      for (int q = P.rowptr[i] + lane; q < P.rowptr[i + 1]; q += 32) {
        const int j = P.colidx[q];
        double x = 0.0;
        for (int a = 0; a < k; ++a) x += U[i + (size_t)n * a] * V[a + (size_t)k * j];
        const double d = x - P.A[i + (size_t)n * j];
        sse += d * d;
      }
    }
    sse = block_sum(sse, red);
    for (int e = warp; e < k * k; e += NW) {  // U'U (kk1), VV' (kk2) for ||UV||_F^2 = <U'U, VV'>
      const int a = e / k, b = e - a * k;
      double s = 0.0, s2 = 0.0;
      for (int i = lane; i < n; i += 32) s += U[i + (size_t)n * a] * U[i + (size_t)n * b];
      for (int j = lane; j < m; j += 32) s2 += V[a + (size_t)k * j] * V[b + (size_t)k * j];
      s = warp_sum(s);
      s2 = warp_sum(s2);
      if (lane == 0) { kk1[e] = s; kk2[e] = s2; }
    }
    __syncthreads();
    double fro = 0.0;
    for (int e = 0; e < k * k; ++e) fro += kk1[e] * kk2[e];
    const double obj = 0.5 * sse + fro / (2.0 * P.gamma);
    if (tid == 0) P.objectives[counter - 1] = obj;
    if (fabs((obj - objective_current) / objective_current) < P.eps) {
      converged = 1;                                                     // OMC.jl:2234-2236
    } else if (counter > 5) {                                            // OMC.jl:2237-2245
      __syncthreads();
      bool all = true;
      for (int q = 0; q < 5; ++q) all = all && (P.objectives[counter - 1 - q] > P.objectives[counter - 6]);
      if (all) converged = 1;
    }
    __syncthreads();
    if (converged) break;
    objective_current = obj;
  }
  if (tid == 0) {
    P.out_int[0] = converged;
    P.out_int[1] = counter;
    P.out_int[2] = inner_total;
  }
}

}  // namespace omc
