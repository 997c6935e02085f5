// omc_big_host.h -- interface between omc_api.cu (the C ABI) and omc_big.cu (the batched large-block relaxation engine).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/omc_b200.h"

namespace omcbig {

struct BigFrontier;
struct ShorHost;

struct BigProblemView {
  int n, m, k, cut_type;
  double gamma, c0;
  const double* AMrm;          // row-major n x m, mask * A        (device)
  const unsigned char* Mkrm;   // row-major n x m, mask            (device)
  const double* pool_x;        // cut pool: [cap][n]               (device)
  const double* pool_vhat;     // cut pool: [cap][k]               (device)
  cudaStream_t stream;
  int sm_count;
  const ShorHost* shor;        // Shor valid-inequality structure of the problem (nullptr: none)
};

struct BigTuning {
  int steps_max, steps_start, seed, infeasible_by_bound, jacobi_sweeps, window;
  double track_tol, confirm_tol;
};

struct BigStats {
  long long launches, node_iterations, rho_changes;
  int iterations, checks;
};

const char* big_last_error();
size_t big_node_bytes(int n, int m, int k, int Lmax);
int big_prepare_problem(int n, int m, const double* A_colmajor, const double* Mk_colmajor, double** AMrm, unsigned char** Mkrm,
                        cudaStream_t st);
void big_default_tuning(BigTuning* t);
int big_create(const BigProblemView& pv, int B, const int* node_cut_ptr, const int* node_cut_ids, const unsigned char* node_cut_dirs,
               BigFrontier** out);
int big_relax(BigFrontier* f, const omc_relax_opts* opts, const BigTuning* tune, float* kernel_ms);
int big_fetch(BigFrontier* f, int* status, double* objective, double* lower_bound, int* iters, double* res, double* X, double* Y,
              double* U);
const BigStats* big_stats(const BigFrontier* f);
// diagnostics: copies one array of a node record to the host.  which: 0..2 V_b, 3..5 Z_b, 6..8 theta_b, 9..11 R_b, 12..14 W_b,
// 15 X, 16 Y, 17 T, 18 U (scaled variables, row-major), 19 the node scalars (rho, v4, rp, rd, objp, objd, lb, ...).  Returns the number of doubles (cap = capacity of out), < 0 on error.
long long big_debug_fetch(BigFrontier* f, int node, int which, double* out, long long cap);
void big_destroy(BigFrontier* f);
// Shor rows (OMC.jl:1503-1552, 1755-1828): minors [nm][4] = (i1, i2, j1, j2) and SOC coordinates [nsoc][2], 0-based host arrays
int big_shor_create(int n, int m, int k, long long nm, const int* minors, long long nsoc, const int* soc, ShorHost** out);
void big_shor_destroy(ShorHost* h);
int big_fetch_shor(BigFrontier* f, double* W, double* Xt);
// separation oracle for n > 104 (restarted Lanczos, one CTA per node); device pointers, column-major Y / U per node
int big_smallest_eigvecs(int n, int k, int B, const double* dY, const double* dU, int nev, double* dlam, double* dvec, double* dbp,
                         int* dfeas, cudaStream_t st);

}  // namespace omcbig
