// omc_big.cu -- host side of the batched large-block relaxation engine (kernels: omc_big.cuh).  Drives the lockstep ADMM
// of a frontier: one short kernel sequence per iteration, a residual check every `check_every` iterations after which
// the host compacts the list of active nodes.  Called from omc_api.cu behind omc_frontier_* (include/omc_b200.h).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <chrono>
#include <string>
#include <vector>

#include <algorithm>
#include <unordered_map>

#include "omc_big.cuh"
#include "omc_big_shor.cuh"
#include "omc_big_host.h"

namespace omcbig {

struct ShorHost {   // problem-level Shor structure on the device (shared by all nodes)
  long long nm = 0, nv1 = 0, nv2 = 0;
  int n = 0, m = 0, k = 0;
  int *minors = nullptr, *mv = nullptr, *cptr = nullptr, *cinc = nullptr, *v1ptr = nullptr, *v1inc = nullptr, *v2ptr = nullptr,
      *v2inc = nullptr, *cnt = nullptr;
  unsigned char* flags = nullptr;
};

struct BigFrontier {
  double* SS = nullptr;        // Shor node records
  ShorLayout SL;
  double *outW = nullptr, *outXt = nullptr;
  BigProblemView pv;
  Layout L;
  int B = 0, E = 0, Lmax = 0;
  double* S = nullptr;
  int* I = nullptr;
  int* active = nullptr;
  int* counters = nullptr;
  int* cut_ptr = nullptr; int* cut_ids = nullptr; unsigned char* cut_dirs = nullptr;
  double* base[3] = {nullptr, nullptr, nullptr};
  double *outX = nullptr, *outY = nullptr, *outU = nullptr, *outT = nullptr, *objective = nullptr, *lower_bound = nullptr, *res = nullptr;
  int *status = nullptr, *iters = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  size_t smem_y = 0, smem_xt = 0, smem_small = 0, smem_init = 0;
  BigStats stats;
  std::vector<int> h_int;
};

static std::string g_berr;
const char* big_last_error() { return g_berr.c_str(); }

#define BCU(call)                                                                                             \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) {                                                                                  \
      char buf_[512];                                                                                         \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      g_berr = buf_;                                                                                          \
      return -2;                                                                                              \
    }                                                                                                         \
  } while (0)

__global__ void k_rowmajor(const double* __restrict__ A, const double* __restrict__ Mk, int n, int m, double* __restrict__ AM,
                           unsigned char* __restrict__ Mb) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)n * m; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e - (size_t)i * m);
    const double mk = Mk[(size_t)j * n + i];
    AM[e] = mk * A[(size_t)j * n + i];
    Mb[e] = mk != 0.0 ? 1 : 0;
  }
}

template <typename T>
static int upload(T** dst, const std::vector<T>& v) {
  const size_t cnt = v.empty() ? 1 : v.size();
  BCU(cudaMalloc(dst, cnt * sizeof(T)));
  if (!v.empty()) BCU(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int big_shor_create(int n, int m, int k, long long nm, const int* minors, long long nsoc, const int* soc, ShorHost** out) {
  if (k > 4) { g_berr = "Shor rows: k > 4 unsupported ((k+1) x (k+1) blocks are projected in registers up to order 5)"; return -4; }
  ShorHost* h = new ShorHost();
  h->n = n; h->m = m; h->k = k; h->nm = nm;
  const size_t C = (size_t)n * m;
  std::vector<unsigned char> flags(C, 0);
  std::vector<int> cnt(C, 0), mvec((size_t)4 * nm), mv((size_t)4 * nm);
  std::unordered_map<unsigned long long, int> id1, id2;
  id1.reserve((size_t)nm); id2.reserve((size_t)nm);
  auto key3 = [](int a_, int b_, int c_) { return ((unsigned long long)a_ << 42) | ((unsigned long long)b_ << 21) | (unsigned long long)c_; };
  for (long long q = 0; q < nm; ++q) {
    const int i1 = minors[4 * q], i2 = minors[4 * q + 1], j1 = minors[4 * q + 2], j2 = minors[4 * q + 3];
    if (i1 < 0 || i1 >= i2 || i2 >= n || j1 < 0 || j1 >= j2 || j2 >= m) { delete h; g_berr = "Shor minor out of range (need 0 <= i1 < i2 < n, 0 <= j1 < j2 < m)"; return -1; }
    for (int e = 0; e < 4; ++e) mvec[4 * q + e] = minors[4 * q + e];
    const size_t cc[4] = {(size_t)i1 * m + j1, (size_t)i1 * m + j2, (size_t)i2 * m + j1, (size_t)i2 * m + j2};
    for (int e = 0; e < 4; ++e) { flags[cc[e]] |= 1; cnt[cc[e]] += 1; }
    auto get = [](std::unordered_map<unsigned long long, int>& mp, unsigned long long key) {
      auto it = mp.find(key);
      if (it != mp.end()) return it->second;
      const int id = (int)mp.size();
      mp.emplace(key, id);
      return id;
    };
    mv[4 * q + 0] = get(id1, key3(i1, j1, j2)); mv[4 * q + 1] = get(id1, key3(i2, j1, j2));
    mv[4 * q + 2] = get(id2, key3(i1, i2, j1)); mv[4 * q + 3] = get(id2, key3(i1, i2, j2));
  }
  for (long long q = 0; q < nsoc; ++q) {
    const int i = soc[2 * q], j = soc[2 * q + 1];
    if (i < 0 || i >= n || j < 0 || j >= m) { delete h; g_berr = "Shor SOC coordinate out of range"; return -1; }
    if (flags[(size_t)i * m + j] & 1) { delete h; g_berr = "a SOC coordinate is covered by a minor (OMC.jl:656-665 lists only uncovered ones)"; return -1; }
    flags[(size_t)i * m + j] |= 2;
  }
  h->nv1 = (long long)id1.size(); h->nv2 = (long long)id2.size();
  // incidence lists (CSR): coordinates <- (minor, slot), V1 ids <- (minor, which), V2 ids <- (minor, which)
  std::vector<int> cptr(C + 1, 0), cinc((size_t)4 * nm), v1ptr(h->nv1 + 1, 0), v1inc((size_t)2 * nm), v2ptr(h->nv2 + 1, 0), v2inc((size_t)2 * nm);
  for (size_t c = 0; c < C; ++c) cptr[c + 1] = cptr[c] + cnt[c];
  {
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (long long q = 0; q < nm; ++q) {
      const int i1 = mvec[4 * q], i2 = mvec[4 * q + 1], j1 = mvec[4 * q + 2], j2 = mvec[4 * q + 3];
      const size_t cc[4] = {(size_t)i1 * m + j1, (size_t)i1 * m + j2, (size_t)i2 * m + j1, (size_t)i2 * m + j2};
      for (int e = 0; e < 4; ++e) cinc[fill[cc[e]]++] = (int)(q * 4 + e);
    }
  }
  for (long long q = 0; q < nm; ++q) { v1ptr[mv[4 * q] + 1]++; v1ptr[mv[4 * q + 1] + 1]++; v2ptr[mv[4 * q + 2] + 1]++; v2ptr[mv[4 * q + 3] + 1]++; }
  for (long long v = 0; v < h->nv1; ++v) v1ptr[v + 1] += v1ptr[v];
  for (long long v = 0; v < h->nv2; ++v) v2ptr[v + 1] += v2ptr[v];
  {
    std::vector<int> f1(v1ptr.begin(), v1ptr.end() - 1), f2(v2ptr.begin(), v2ptr.end() - 1);
    for (long long q = 0; q < nm; ++q) {
      v1inc[f1[mv[4 * q]]++] = (int)(q * 2); v1inc[f1[mv[4 * q + 1]]++] = (int)(q * 2 + 1);
      v2inc[f2[mv[4 * q + 2]]++] = (int)(q * 2); v2inc[f2[mv[4 * q + 3]]++] = (int)(q * 2 + 1);
    }
  }
  if (upload(&h->minors, mvec) || upload(&h->mv, mv) || upload(&h->cptr, cptr) || upload(&h->cinc, cinc) || upload(&h->v1ptr, v1ptr) ||
      upload(&h->v1inc, v1inc) || upload(&h->v2ptr, v2ptr) || upload(&h->v2inc, v2inc) || upload(&h->cnt, cnt) || upload(&h->flags, flags)) {
    delete h;
    return -2;
  }
  *out = h;
  return 0;
}

void big_shor_destroy(ShorHost* h) {
  if (!h) return;
  cudaFree(h->minors); cudaFree(h->mv); cudaFree(h->cptr); cudaFree(h->cinc); cudaFree(h->v1ptr); cudaFree(h->v1inc);
  cudaFree(h->v2ptr); cudaFree(h->v2inc); cudaFree(h->cnt); cudaFree(h->flags);
  delete h;
}

size_t big_node_bytes(int n, int m, int k, int Lmax) { return make_layout(n, m, k, Lmax > 0 ? Lmax : 0).total * sizeof(double); }

int big_create(const BigProblemView& pv, int B, const int* node_cut_ptr, const int* node_cut_ids, const unsigned char* node_cut_dirs,
               BigFrontier** out) {
  BigFrontier* f = new BigFrontier();
  f->pv = pv; f->B = B; f->E = node_cut_ptr[B];
  int Lmax = 0;
  for (int b = 0; b < B; ++b) Lmax = std::max(Lmax, node_cut_ptr[b + 1] - node_cut_ptr[b]);
  f->Lmax = Lmax;
  f->L = make_layout(pv.n, pv.m, pv.k, Lmax);
  const Layout& L = f->L;
  if (pv.k > PM) { g_berr = "large-block engine: k > 16 unsupported"; delete f; return -4; }
  f->smem_y = ((size_t)YOFF_X + 2 * (size_t)L.Lcap * TS + L.Lcap + 1 + 8 * (size_t)L.rcap + 32) * sizeof(double);
  f->smem_xt = ((size_t)2 * TS * ZLD + PM + 8) * sizeof(double);
  static_assert(2 * TS * ZLD >= TS * (TS + 1) && YOFF_SC >= TS * (TS + 1), "the transposition buffer aliases the panels");
  f->smem_small = (size_t)4 * L.rcap * sizeof(double);
  f->smem_init = (size_t)(Lmax * Lmax + 1) * sizeof(double);
  if (f->smem_y > 227 * 1024 || f->smem_init > 227 * 1024) {
    char buf[256];
    snprintf(buf, sizeof buf, "large-block engine: a node with %d cuts needs %zu bytes of shared memory (limit 227 KB, about 100 cuts)", Lmax, f->smem_y);
    g_berr = buf; delete f; return -4;
  }
  cudaStream_t st = pv.stream;
  size_t free_b = 0, tot_b = 0;
  BCU(cudaMemGetInfo(&free_b, &tot_b));
  const size_t need = (size_t)B * L.total * sizeof(double);
  if (need + ((size_t)1 << 30) > free_b) {
    char buf[256];
    snprintf(buf, sizeof buf, "large-block engine: %d nodes x %.1f MB exceed the free device memory (%.1f GB); relax the frontier in chunks",
             B, L.total * 8.0 / 1048576.0, free_b / 1073741824.0);
    g_berr = buf; delete f; return -4;
  }
  if (pv.shor) {
    f->SL = make_shor_layout(pv.n, pv.m, pv.k, pv.shor->nm, pv.shor->nv1, pv.shor->nv2);
    const size_t need_s = (size_t)B * f->SL.total * sizeof(double);
    if (need + need_s + ((size_t)1 << 30) > free_b) {
      char buf[256];
      snprintf(buf, sizeof buf, "Shor rows: %d nodes x %.1f MB exceed the free device memory; relax the frontier in chunks", B, f->SL.total * 8.0 / 1048576.0);
      g_berr = buf; delete f; return -4;
    }
    BCU(cudaMalloc(&f->SS, need_s));
    BCU(cudaMalloc(&f->outW, (size_t)B * pv.n * pv.m * sizeof(double)));
    BCU(cudaMalloc(&f->outXt, (size_t)B * pv.k * pv.n * pv.m * sizeof(double)));
  }
  BCU(cudaMalloc(&f->S, need));
  BCU(cudaMalloc(&f->I, (size_t)B * ISTR * sizeof(int)));
  BCU(cudaMalloc(&f->active, (size_t)B * sizeof(int)));
  BCU(cudaMalloc(&f->counters, 8 * sizeof(int)));
  BCU(cudaMalloc(&f->cut_ptr, (size_t)(B + 1) * sizeof(int)));
  BCU(cudaMalloc(&f->cut_ids, (size_t)std::max(f->E, 1) * sizeof(int)));
  BCU(cudaMalloc(&f->cut_dirs, (size_t)std::max(f->E, 1) * pv.k));
  for (int b = 0; b < 3; ++b) BCU(cudaMalloc(&f->base[b], (size_t)L.N[b] * PM * sizeof(double)));
  BCU(cudaMalloc(&f->outX, (size_t)B * pv.n * pv.m * sizeof(double)));
  BCU(cudaMalloc(&f->outY, (size_t)B * pv.n * pv.n * sizeof(double)));
  BCU(cudaMalloc(&f->outU, (size_t)B * pv.n * pv.k * sizeof(double)));
  BCU(cudaMalloc(&f->objective, (size_t)B * sizeof(double)));
  BCU(cudaMalloc(&f->lower_bound, (size_t)B * sizeof(double)));
  BCU(cudaMalloc(&f->res, (size_t)2 * B * sizeof(double)));
  BCU(cudaMalloc(&f->status, (size_t)B * sizeof(int)));
  BCU(cudaMalloc(&f->iters, (size_t)B * sizeof(int)));
  BCU(cudaMemcpyAsync(f->cut_ptr, node_cut_ptr, (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (f->E > 0) {
    BCU(cudaMemcpyAsync(f->cut_ids, node_cut_ids, (size_t)f->E * sizeof(int), cudaMemcpyHostToDevice, st));
    BCU(cudaMemcpyAsync(f->cut_dirs, node_cut_dirs, (size_t)f->E * pv.k, cudaMemcpyHostToDevice, st));
  }
  BCU(cudaEventCreate(&f->ev0));
  BCU(cudaEventCreate(&f->ev1));
  BCU(cudaFuncSetAttribute(k_y1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_y));
  BCU(cudaFuncSetAttribute(k_y2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_y));
  BCU(cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_y));
  BCU(cudaFuncSetAttribute(k_xt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_xt));
  BCU(cudaFuncSetAttribute(k_node_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_init));
  BCU(cudaFuncSetAttribute(k_rr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM));
  BCU(cudaFuncSetAttribute(k_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM));
  BCU(cudaStreamSynchronize(st));
  f->h_int.resize((size_t)B * ISTR);
  *out = f;
  return 0;
}

void big_destroy(BigFrontier* f) {
  if (!f) return;
  cudaFree(f->SS); cudaFree(f->outW); cudaFree(f->outXt);
  cudaFree(f->S); cudaFree(f->I); cudaFree(f->active); cudaFree(f->counters);
  cudaFree(f->cut_ptr); cudaFree(f->cut_ids); cudaFree(f->cut_dirs);
  for (int b = 0; b < 3; ++b) cudaFree(f->base[b]);
  cudaFree(f->outX); cudaFree(f->outY); cudaFree(f->outU); cudaFree(f->outT);
  cudaFree(f->objective); cudaFree(f->lower_bound); cudaFree(f->res); cudaFree(f->status); cudaFree(f->iters);
  if (f->ev0) cudaEventDestroy(f->ev0);
  if (f->ev1) cudaEventDestroy(f->ev1);
  delete f;
}

int big_relax(BigFrontier* f, const omc_relax_opts* ro, const BigTuning* tune, float* kernel_ms) {
  const BigProblemView& pv = f->pv;
  const Layout& L = f->L;
  cudaStream_t st = pv.stream;
  const int B = f->B;
  BigArgs a;
  memset(&a, 0, sizeof a);
  a.L = L; a.S = f->S; a.I = f->I; a.active = f->active; a.AM = pv.AMrm; a.Mk = pv.Mkrm;
  a.a = (double)pv.n; a.sa = sqrt((double)pv.n); a.cT = a.a / (2.0 * pv.gamma); a.c0 = pv.c0; a.ktr = a.a * pv.k;
  a.o.eps_abs = ro->eps_abs; a.o.eps_rel = ro->eps_rel; a.o.sigma = ro->sigma; a.o.alpha = ro->alpha; a.o.rho0 = ro->rho0;
  a.o.cutoff = ro->cutoff; a.o.max_iter = ro->max_iter; a.o.check_every = ro->check_every; a.o.adapt_every = ro->adapt_every;
  a.o.fix_linear3_right = ro->fix_linear3_right; a.o.cut_type = pv.cut_type;
  a.o.track_tol = tune->track_tol; a.o.confirm_tol = tune->confirm_tol; a.o.adapt_thresh = 5.0;
  a.o.jacobi_sweeps = tune->jacobi_sweeps;
  a.o.window = tune->window;
  a.o.steps_max = tune->steps_max; a.o.steps_start = tune->steps_start; a.o.infeasible_by_bound = tune->infeasible_by_bound;
  const bool shor = pv.shor != nullptr;
  if (shor) {
    const ShorHost* h = pv.shor;
    a.sh.on = 1; a.sh.SL = f->SL; a.sh.SS = f->SS; a.sh.minors = h->minors; a.sh.mv = h->mv; a.sh.cptr = h->cptr; a.sh.cinc = h->cinc;
    a.sh.v1ptr = h->v1ptr; a.sh.v1inc = h->v1inc; a.sh.v2ptr = h->v2ptr; a.sh.v2inc = h->v2inc; a.sh.flags = h->flags; a.sh.cnt = h->cnt;
  }
  const ShorLayout& SL = f->SL;
  const long long nblk5 = shor ? (long long)SL.k * SL.nm : 0, nvv = shor ? SL.nv1 + SL.nv2 + SL.nm : 0;
  const unsigned g5 = (unsigned)std::max<long long>(1, (nblk5 + 127) / 128), gc128 = shor ? (unsigned)((SL.C + 127) / 128) : 1,
                 gc256 = shor ? (unsigned)((SL.C + 255) / 256) : 1, gcw = shor ? (unsigned)((SL.C + 7) / 8) : 1,
                 gv = (unsigned)std::max<long long>(1, (nvv + 255) / 256), gvr = (unsigned)std::max<long long>(1, ((long long)SL.k * nvv + 255) / 256);
  const auto t_start = std::chrono::steady_clock::now();
  BCU(cudaEventRecord(f->ev0, st));
  // ---- setup: start bases, node records, Woodbury inverses
  k_start_basis<<<3, 256, 0, st>>>(L, f->base[0], f->base[1], f->base[2], tune->seed);
  BCU(cudaMemsetAsync(f->S, 0, (size_t)B * L.total * sizeof(double), st));
  SetupArgs sa;
  sa.pool_x = pv.pool_x; sa.pool_vhat = pv.pool_vhat; sa.cut_ptr = f->cut_ptr; sa.cut_ids = f->cut_ids; sa.cut_dirs = f->cut_dirs;
  for (int b = 0; b < 3; ++b) sa.base[b] = f->base[b];
  k_node_init<<<B, 256, f->smem_init, st>>>(a, sa);
  if (shor) {
    BCU(cudaMemsetAsync(f->SS, 0, (size_t)B * SL.total * sizeof(double), st));
    k_shor_init<<<dim3(std::max(g5, gc256), B), 256, 0, st>>>(a);
  }
  std::vector<int> act(B);
  for (int b = 0; b < B; ++b) act[b] = b;
  BCU(cudaMemcpyAsync(f->active, act.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice, st));
  int nact = B;
  k_minv<<<nact, 256, 0, st>>>(a, 0);
  BCU(cudaGetLastError());
  int ntmax = std::max(L.nt[0], std::max(L.nt[1], L.nt[2]));
  int it = 0, force = 0;
  int h_cnt[8];
  BigStats& stt = f->stats;
  memset(&stt, 0, sizeof stt);
  bool timed_out = false;
  while (nact > 0 && it < a.o.max_iter) {
    ++it;
    a.it = it; a.force = force; a.step = 0;
    if (shor) {     // Shor rows: projections of the moment blocks, adjoint sums into the variables (omc_big_shor.cuh)
      k_shor_proj5<<<dim3(g5, nact), 128, 0, st>>>(a, 0);
      k_shor_proj9<<<dim3(gc128, nact), 128, 0, st>>>(a, 0);
      k_shor_gather_c<<<dim3(gcw, nact), 256, 0, st>>>(a, 0);
      k_shor_gather_v<<<dim3(gv, nact), 256, 0, st>>>(a, 0);
      stt.launches += 4;
    }
    k_xt<<<dim3(L.tilesXT, nact), 256, f->smem_xt, st>>>(a);
    if (shor) { k_shor_col<<<dim3(L.m, nact), 128, 0, st>>>(a); stt.launches += 1; }
    k_y1<<<dim3(L.tilesYU, nact), 256, f->smem_y, st>>>(a);
    k_small<<<nact, 128, f->smem_small, st>>>(a);
    k_y2<<<dim3(L.tilesYU, nact), 256, f->smem_y, st>>>(a);
    if (shor) {
      k_shor_v5<<<dim3(g5, nact), 128, 0, st>>>(a);
      k_shor_vc<<<dim3(gc256, nact), 256, 0, st>>>(a);
      k_shor_relax_v<<<dim3(gvr, nact), 256, 0, st>>>(a);
      stt.launches += 3;
    }
    const bool check = (it % a.o.check_every == 0) || it >= a.o.max_iter;
    // steps_start rounds at the first iteration and, when some node has a termination decision pending, in check iterations
    const int rounds = (it == 1 || (check && force) || it >= a.o.max_iter) ? a.o.steps_start : a.o.steps_max;
    for (int s = 0; s < rounds; ++s) {
      a.step = s;
      if (s == 0) k_prod<<<dim3(ntmax, nact, 3), 128, 0, st>>>(a, 0);
      k_resid<<<dim3(ntmax, nact, 3), 256, 0, st>>>(a, 0);
      k_resid<<<dim3(ntmax, nact, 3), 256, 0, st>>>(a, 1);
      k_resid<<<dim3(ntmax, nact, 3), 256, 0, st>>>(a, 2);
      k_gram<<<dim3(nact, 3), 256, 0, st>>>(a);
      k_prod<<<dim3(ntmax, nact, 3), 128, 0, st>>>(a, 1);
      k_rr<<<dim3(nact, 3), 256, RR_SMEM, st>>>(a);
      k_update<<<dim3(ntmax, nact, 3), 256, UPD_SMEM, st>>>(a);
      stt.launches += (s == 0) ? 8 : 7;
    }
    stt.launches += 4;
    stt.node_iterations += nact;
    if (!check) continue;
    k_reorth<<<dim3(nact, 3), 256, 0, st>>>(a);
    BCU(cudaMemsetAsync(f->counters, 0, 8 * sizeof(int), st));
    if (shor) {     // multipliers mu / rho = v - P(v) of the Shor rows and their adjoint sums
      k_shor_zero_chk<<<nact, 32, 0, st>>>(a);
      k_shor_proj5<<<dim3(g5, nact), 128, 0, st>>>(a, 1);
      k_shor_proj9<<<dim3(gc128, nact), 128, 0, st>>>(a, 1);
      k_shor_gather_c<<<dim3(gcw, nact), 256, 0, st>>>(a, 1);
      k_shor_gather_v<<<dim3(gv, nact), 256, 0, st>>>(a, 1);
    }
    k_check<<<dim3(L.tilesAll, nact), 256, f->smem_y, st>>>(a);
    if (shor) {
      k_shor_check_c<<<dim3(gc256, nact), 256, 0, st>>>(a);
      k_shor_check_5<<<dim3(g5, nact), 128, 0, st>>>(a);
      k_shor_check_col<<<dim3((L.m + 127) / 128, nact), 128, 0, st>>>(a);
      stt.launches += 8;
    }
    k_decide<<<nact, 128, (size_t)L.rcap * sizeof(double), st>>>(a, f->counters);
    stt.launches += 3; stt.checks += 1;
    BCU(cudaMemcpyAsync(h_cnt, f->counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    BCU(cudaStreamSynchronize(st));
    BCU(cudaGetLastError());
    if (h_cnt[2] > 0) {
      if (shor) {
        k_shor_rescale5<<<dim3(g5, nact), 128, 0, st>>>(a);
        k_shor_rescale_c<<<dim3(gc256, nact), 256, 0, st>>>(a);
      }
      k_rescale<<<dim3(ntmax, nact, 4), 256, 0, st>>>(a);
      k_rescale_theta<<<nact, 3 * PM, 0, st>>>(a);
      k_minv<<<nact, 256, 0, st>>>(a, 1);
      stt.launches += 3; stt.rho_changes += h_cnt[2];
    }
    force = h_cnt[1] > 0 ? 1 : 0;
    if (h_cnt[0] != nact) {   // some nodes finished: compact the active list
      BCU(cudaMemcpyAsync(f->h_int.data(), f->I, (size_t)B * ISTR * sizeof(int), cudaMemcpyDeviceToHost, st));
      BCU(cudaStreamSynchronize(st));
      int q = 0;
      for (int b = 0; b < B; ++b) if (!f->h_int[(size_t)b * ISTR + I_DONE]) act[q++] = b;
      nact = q;
      if (nact > 0) BCU(cudaMemcpyAsync(f->active, act.data(), (size_t)nact * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (ro->time_limit_s > 0.0 && nact > 0) {
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
      if (el > ro->time_limit_s) { timed_out = true; break; }
    }
  }
  stt.iterations = it;
  k_extract<<<B, 256, 0, st>>>(a, B, f->outX, f->outY, f->outU, nullptr, f->status, f->iters, f->objective, f->lower_bound, f->res);
  if (shor) k_shor_extract<<<dim3(gc256, B), 256, 0, st>>>(a, B, f->outW, f->outXt);
  BCU(cudaGetLastError());
  BCU(cudaEventRecord(f->ev1, st));
  BCU(cudaStreamSynchronize(st));
  if (timed_out) {   // nodes still active: MOI.TIME_LIMIT with the values of their last check
    std::vector<int> hs(B);
    BCU(cudaMemcpy(hs.data(), f->status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b) if (hs[b] < 0) hs[b] = OMC_STATUS_TIME_LIMIT;
    BCU(cudaMemcpy(f->status, hs.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (kernel_ms) BCU(cudaEventElapsedTime(kernel_ms, f->ev0, f->ev1));
  return 0;
}

int big_fetch(BigFrontier* f, int* status, double* objective, double* lower_bound, int* iters, double* res, double* X, double* Y,
              double* U) {
  const BigProblemView& pv = f->pv;
  cudaStream_t st = pv.stream;
  const size_t B = f->B;
  if (status) BCU(cudaMemcpyAsync(status, f->status, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (iters) BCU(cudaMemcpyAsync(iters, f->iters, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (objective) BCU(cudaMemcpyAsync(objective, f->objective, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (lower_bound) BCU(cudaMemcpyAsync(lower_bound, f->lower_bound, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (res) BCU(cudaMemcpyAsync(res, f->res, 2 * B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (X) BCU(cudaMemcpyAsync(X, f->outX, B * pv.n * pv.m * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (Y) BCU(cudaMemcpyAsync(Y, f->outY, B * pv.n * pv.n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (U) BCU(cudaMemcpyAsync(U, f->outU, B * pv.n * pv.k * sizeof(double), cudaMemcpyDeviceToHost, st));
  BCU(cudaStreamSynchronize(st));
  return 0;
}

const BigStats* big_stats(const BigFrontier* f) { return &f->stats; }

int big_fetch_shor(BigFrontier* f, double* W, double* Xt) {
  if (!f->SS) { g_berr = "the frontier has no Shor rows"; return -3; }
  const BigProblemView& pv = f->pv;
  const size_t B = f->B, C = (size_t)pv.n * pv.m;
  if (W) BCU(cudaMemcpyAsync(W, f->outW, B * C * sizeof(double), cudaMemcpyDeviceToHost, pv.stream));
  if (Xt) BCU(cudaMemcpyAsync(Xt, f->outXt, B * pv.k * C * sizeof(double), cudaMemcpyDeviceToHost, pv.stream));
  BCU(cudaStreamSynchronize(pv.stream));
  return 0;
}

long long big_debug_fetch(BigFrontier* f, int node, int which, double* out, long long cap) {
  const Layout& L = f->L;
  if (node < 0 || node >= f->B || which < 0 || which > 19) return -1;
  size_t off = 0, cnt = 0;
  const int b = which % 3;
  if (which < 3) { off = L.V[b]; cnt = (size_t)L.N[b] * L.N[b]; }
  else if (which < 6) { off = L.Z[b]; cnt = (size_t)L.N[b] * PM; }
  else if (which < 9) { off = L.th[b]; cnt = PM; }
  else if (which < 12) { off = L.R[b]; cnt = (size_t)L.N[b] * PM; }
  else if (which < 15) { off = L.W[b]; cnt = (size_t)L.N[b] * PM; }
  else if (which == 15) { off = L.X; cnt = (size_t)L.n * L.m; }
  else if (which == 16) { off = L.Y; cnt = (size_t)L.n * L.n; }
  else if (which == 17) { off = L.T; cnt = (size_t)L.m * L.m; }
  else if (which == 19) { off = L.scal; cnt = SSTR; }
  else { off = L.U; cnt = (size_t)L.n * L.k; }
  if ((long long)cnt > cap) return -1;
  if (cudaMemcpy(out, f->S + (size_t)node * L.total + off, cnt * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  return (long long)cnt;
}

int big_prepare_problem(int n, int m, const double* A, const double* Mk, double** AMrm, unsigned char** Mkrm, cudaStream_t st) {
  BCU(cudaMalloc(AMrm, (size_t)n * m * sizeof(double)));
  BCU(cudaMalloc(Mkrm, (size_t)n * m));
  k_rowmajor<<<148 * 4, 256, 0, st>>>(A, Mk, n, m, *AMrm, *Mkrm);
  BCU(cudaGetLastError());
  BCU(cudaStreamSynchronize(st));
  return 0;
}

int big_smallest_eigvecs(int n, int k, int B, const double* dY, const double* dU, int nev, double* dlam, double* dvec, double* dbp,
                         int* dfeas, cudaStream_t st) {
  if (k > 16) { g_berr = "separation oracle: k > 16 unsupported"; return -4; }
  double* ws = nullptr;
  BCU(cudaMalloc(&ws, (size_t)B * (LM + 3) * n * sizeof(double)));
  const size_t smem = LANCZOS_SMEM_FIXED + (size_t)n * sizeof(double);
  if (smem > 227 * 1024) { cudaFree(ws); g_berr = "separation oracle: n too large for the shared-memory vector"; return -4; }
  BCU(cudaFuncSetAttribute(k_lanczos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_lanczos<<<B, 256, smem, st>>>(n, k, B, dY, dU, nev, ws, dlam, dvec, dbp, dfeas, 1e-10, 40);
  BCU(cudaGetLastError());
  BCU(cudaStreamSynchronize(st));
  cudaFree(ws);
  return 0;
}

void big_default_tuning(BigTuning* t) {
  t->jacobi_sweeps = 3;
  t->window = 2;
  t->steps_max = 3; t->steps_start = 6; t->track_tol = 1e-3; t->confirm_tol = 1e-9; t->seed = 1; t->infeasible_by_bound = 1;
}

}  // namespace omcbig
