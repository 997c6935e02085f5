// omc_eigsep.cuh -- batched separation oracle (K5): the smallest one or two eigenpairs of U U' - Y.
// Replaces eigs(Symmetric(U*U' - Y), nev, which=:SR, tol=1e-6) at OMC.jl:2466-2477 and the
// feasibility test lambda_min >= -1e-6 at OMC.jl:1272-1277.  One CTA per node: the n x n matrix is
// assembled in shared memory (U U' is never stored in HBM) and diagonalised by the same parallel
// Jacobi eigensolver the relaxation kernel uses (block reductions by warp shuffles), so the result is
// accurate to rounding -- tighter than ARPACK's tol = 1e-6 -- and deterministic: ARPACK's random
// start vector is replaced by the sign rule "largest-|.| component positive".
#pragma once
#include "omc_device.cuh"

namespace omc {

template <int NT>
__global__ void __launch_bounds__(NT, 1) eigsep_kernel(int n, int k, int B, const double* __restrict__ Y,
                                                      const double* __restrict__ U, int nev, double* __restrict__ lam_out,
                                                      double* __restrict__ vec_out, double* __restrict__ bp_out,
                                                      int* __restrict__ feas_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Geo g = make_geo(n);
  const int NP = g.NP, ld = g.ld, tid = threadIdx.x;
  double* S = reinterpret_cast<double*>(smem_raw);
  double* Q = S + (size_t)NP * ld;
  double* lam = Q + (size_t)NP * ld;
  double* jcs = lam + NP;
  double* jsn = jcs + NP / 2;
  double* red = jsn + NP / 2;
  double* wsm = red + 32;  // [2] mixing weights
  int* jrot = reinterpret_cast<int*>(wsm + 2);
  int* sel = jrot + 3 * (NP / 2);  // [4]: index of smallest, second smallest, sign flips
  for (int node = blockIdx.x; node < B; node += gridDim.x) {
    const double* Yn = Y + (size_t)node * n * n;  // column-major
    const double* Un = U + (size_t)node * n * k;  // column-major
    __syncthreads();
    for (int e = tid; e < NP * NP; e += NT) {
      const int r = e / NP, c = e - r * NP;
      double v = 0.0;
      if (r < n && c < n) {
        for (int t = 0; t < k; ++t) v += Un[r + (size_t)n * t] * Un[c + (size_t)n * t];
        v -= 0.5 * (Yn[r + (size_t)n * c] + Yn[c + (size_t)n * r]);
      }
      S[(size_t)r * ld + c] = v;
      Q[(size_t)r * ld + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    jacobi_sym(S, Q, NP, ld, 1e-13, 60, jcs, jsn, jrot, red);
    for (int i = tid; i < NP; i += NT) lam[i] = S[(size_t)i * ld + i];
    __syncthreads();
    if (tid == 0) {
      int i0 = 0, i1 = -1;
      for (int i = 1; i < n; ++i)
        if (lam[i] < lam[i0]) i0 = i;
      for (int i = 0; i < n; ++i)
        if (i != i0 && (i1 < 0 || lam[i] < lam[i1])) i1 = i;
      sel[0] = i0;
      sel[1] = (i1 < 0) ? i0 : i1;
      for (int q = 0; q < 2; ++q) {  // sign: largest-|.| component positive, first index on ties
        const int col = sel[q];
        int best = 0;
        double bv = fabs(Q[col]);
        for (int i = 1; i < n; ++i) {
          const double a = fabs(Q[(size_t)i * ld + col]);
          if (a > bv) { bv = a; best = i; }
        }
        sel[2 + q] = (Q[(size_t)best * ld + col] < 0.0) ? -1 : 1;
      }
      const double l0 = lam[i0], l1 = lam[sel[1]];
      if (nev == 2 && l1 < -1e-10) {  // OMC.jl:2471-2473
        const double nr = sqrt(l0 * l0 + l1 * l1);
        wsm[0] = fabs(l0) / nr;
        wsm[1] = fabs(l1) / nr;
      } else {  // OMC.jl:2468, 2475
        wsm[0] = 1.0;
        wsm[1] = 0.0;
      }
      for (int q = 0; q < nev; ++q) lam_out[(size_t)node * nev + q] = lam[sel[q]];
      feas_out[node] = (l0 >= -1e-6) ? 1 : 0;  // OMC.jl:1274-1276
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      const double v0 = sel[2] * Q[(size_t)i * ld + sel[0]];
      const double v1 = sel[3] * Q[(size_t)i * ld + sel[1]];
      vec_out[(size_t)node * n * nev + i] = v0;
      if (nev == 2) vec_out[(size_t)node * n * nev + n + i] = v1;
      bp_out[(size_t)node * n + i] = wsm[0] * v0 + wsm[1] * v1;
    }
  }
}

inline size_t eigsep_smem_bytes(int n) {
  const Geo g = make_geo(n);
  return ((size_t)2 * g.NP * g.ld + 2 * g.NP + 32 + 2) * 8 + (3 * ((size_t)g.NP / 2) + 4) * 4 + 128;
}

}  // namespace omc
