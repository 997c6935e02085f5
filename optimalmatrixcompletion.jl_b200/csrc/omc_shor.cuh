// omc_shor.cuh -- Shor 2x2-minor index enumeration on the GPU (bit-exact integer work), replacing
// generate_rank1_matrix_completion_Shor_constraints_indexes (/root/reference/src/OptimalMatrixCompletion.jl:2545-2612)
// and the SOC coordinate list built from it (OMC.jl:656-665).
//
// Order contract (what Julia's push! sequence produces): for each num_entries_present p of the list IN THE ORDER GIVEN, one
// or two *phases*; inside a phase the row pairs (i1 < i2) in lexicographic order; inside a row pair the column pairs as the
// reference's loops visit them.  With B / X / N the ascending lists of columns where both / exactly one / none of the two
// rows are observed:
//   p = 4: combinations(B, 2)             p = 3: B x X (sorted pair)        p = 2: phase (a) B x N (sorted), phase (b) combinations(X, 2)
//   p = 1: X x N (sorted)                 p = 0: combinations(N, 2)
// Pass 1 counts the tuples of every (phase, row pair); the host turns the counts into offsets (exclusive scan, the counts
// are a few thousand integers); pass 2 lets one warp per (phase, row pair) write its tuples at their final positions.
// Tuples and coordinates are 0-based inside the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace omc {

constexpr int SHOR_MAXW = 32;   // row masks of up to 64 * 32 = 2048 columns

__global__ void shor_rowmask_kernel(const double* __restrict__ Mk, int n, int m, int W, unsigned long long* __restrict__ rowbits) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * W) return;
  const int i = e / W, w = e - i * W;
  unsigned long long bits = 0ull;
  for (int b = 0; b < 64; ++b) {
    const int j = w * 64 + b;
    if (j < m && Mk[(size_t)i + (size_t)n * j] != 0.0) bits |= (1ull << b);
  }
  rowbits[e] = bits;
}

// row pair q (lexicographic over i1 < i2) -> (i1, i2)
__device__ __forceinline__ void shor_unrank_pair(long long q, int n, int& i1, int& i2) {
  // number of pairs with first index < i: i * (2n - i - 1) / 2
  int i = (int)(((2.0 * n - 1.0) - sqrt((2.0 * n - 1.0) * (2.0 * n - 1.0) - 8.0 * (double)q)) * 0.5);
  if (i < 0) i = 0;
  while (i > 0 && (long long)i * (2 * n - i - 1) / 2 > q) --i;
  while ((long long)(i + 1) * (2 * n - i - 2) / 2 <= q) ++i;
  i1 = i;
  i2 = i + 1 + (int)(q - (long long)i * (2 * n - i - 1) / 2);
}

// phase code: 0: p=4, 1: p=3, 2: p=2 (a), 3: p=2 (b), 4: p=1, 5: p=0
__device__ __forceinline__ long long shor_phase_count(int code, long long nb, long long nx, long long nn) {
  switch (code) {
    case 0: return nb * (nb - 1) / 2;
    case 1: return nb * nx;
    case 2: return nb * nn;
    case 3: return nx * (nx - 1) / 2;
    case 4: return nx * nn;
    default: return nn * (nn - 1) / 2;
  }
}

__global__ void shor_count_kernel(const unsigned long long* __restrict__ rowbits, int n, int m, int W, const int* __restrict__ phases,
                                  int nphase, long long npairs, long long* __restrict__ counts) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npairs) return;
  int i1, i2;
  shor_unrank_pair(q, n, i1, i2);
  long long nb = 0, nx = 0;
  for (int w = 0; w < W; ++w) {
    const unsigned long long a = rowbits[(size_t)i1 * W + w], b = rowbits[(size_t)i2 * W + w];
    nb += __popcll(a & b);
    nx += __popcll(a ^ b);
  }
  const long long nn = (long long)m - nb - nx;
  for (int ph = 0; ph < nphase; ++ph) counts[(size_t)ph * npairs + q] = shor_phase_count(phases[ph], nb, nx, nn);
}

// one warp per (phase, row pair); lists of the pair in shared memory (3 * mcap ints per warp)
__global__ void shor_fill_kernel(const unsigned long long* __restrict__ rowbits, int n, int m, int W, const int* __restrict__ phases,
                                 int nphase, long long npairs, const long long* __restrict__ offsets, int* __restrict__ tuples,
                                 unsigned int* __restrict__ covered, int mcap) {
  extern __shared__ int shor_sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  int* LB = shor_sm + (size_t)wib * 3 * mcap;
  int* LX = LB + mcap;
  int* LN = LX + mcap;
  const long long total = (long long)nphase * npairs;
  for (long long item = (long long)blockIdx.x * wpb + wib; item < total; item += (long long)gridDim.x * wpb) {
    const int ph = (int)(item / npairs);
    const long long q = item - (long long)ph * npairs;
    int i1, i2;
    shor_unrank_pair(q, n, i1, i2);
    int nb = 0, nx = 0, nn = 0;
    for (int base = 0; base < m; base += 32) {
      const int j = base + lane;
      int cls = -1;
      if (j < m) {
        const unsigned long long a = rowbits[(size_t)i1 * W + (j >> 6)] >> (j & 63), b = rowbits[(size_t)i2 * W + (j >> 6)] >> (j & 63);
        cls = (int)(a & 1ull) + (int)(b & 1ull);     // 2: both, 1: one, 0: none
      }
      const unsigned mb = __ballot_sync(0xffffffffu, cls == 2), mx = __ballot_sync(0xffffffffu, cls == 1), mn = __ballot_sync(0xffffffffu, cls == 0);
      const unsigned below = (1u << lane) - 1u;
      if (cls == 2) LB[nb + __popc(mb & below)] = j;
      if (cls == 1) LX[nx + __popc(mx & below)] = j;
      if (cls == 0) LN[nn + __popc(mn & below)] = j;
      nb += __popc(mb); nx += __popc(mx); nn += __popc(mn);
    }
    __syncwarp();
    const int code = phases[ph];
    int* out = tuples + 4 * offsets[(size_t)ph * npairs + q];
    const int* L1 = (code == 0 || code == 1 || code == 2) ? LB : ((code == 3 || code == 4) ? LX : LN);
    const int n1 = (code == 0 || code == 1 || code == 2) ? nb : ((code == 3 || code == 4) ? nx : nn);
    if (code == 0 || code == 3 || code == 5) {            // combinations(L1, 2), lexicographic
      for (int a = lane; a < n1; a += 32) {
        const long long base = (long long)a * (2 * n1 - a - 1) / 2;
        for (int b = a + 1; b < n1; ++b) {
          int* t = out + 4 * (base + (b - a - 1));
          t[0] = i1; t[1] = i2; t[2] = L1[a]; t[3] = L1[b];
        }
      }
    } else {                                               // L1 x L2, first factor slowest, pair sorted
      const int* L2 = (code == 1) ? LX : LN;
      const int n2 = (code == 1) ? nx : nn;
      const long long cnt = (long long)n1 * n2;
      for (long long t_ = lane; t_ < cnt; t_ += 32) {
        const int a = (int)(t_ / n2), b = (int)(t_ - (long long)a * n2);
        const int j1 = L1[a], j2 = L2[b];
        int* t = out + 4 * t_;
        t[0] = i1; t[1] = i2; t[2] = j1 < j2 ? j1 : j2; t[3] = j1 < j2 ? j2 : j1;
      }
    }
    // coverage bitmap over the column-major coordinate index i + n j (for the SOC coordinate list)
    if (covered) {
      const long long cntc = shor_phase_count(code, nb, nx, nn);
      if (cntc > 0) {
        const bool useB = (code <= 2), useX = (code == 1 || code == 3 || code == 4), useN = (code == 2 || code == 4 || code == 5);
        // a column of a used list is covered when the tuple count is positive and, for combinations, the list has >= 2 entries
        for (int s_ = 0; s_ < 3; ++s_) {
          const int* Ls = s_ == 0 ? LB : (s_ == 1 ? LX : LN);
          const int ns = s_ == 0 ? nb : (s_ == 1 ? nx : nn);
          const bool used = s_ == 0 ? useB : (s_ == 1 ? useX : useN);
          if (!used) continue;
          for (int a = lane; a < ns; a += 32) {
            const long long c1 = (long long)i1 + (long long)n * Ls[a], c2 = (long long)i2 + (long long)n * Ls[a];
            atomicOr(&covered[c1 >> 5], 1u << (c1 & 31));
            atomicOr(&covered[c2 >> 5], 1u << (c2 & 31));
          }
        }
      }
    }
    __syncwarp();
  }
}

// SOC coordinates = column-major (i fastest) list of the coordinates not covered by any minor (OMC.jl:656-665): one block scan
__global__ void shor_soc_kernel(const unsigned int* __restrict__ covered, long long total, int n, int* __restrict__ soc, long long* __restrict__ nsoc) {
  __shared__ long long base;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (long long c0 = 0; c0 < total; c0 += blockDim.x) {
    const long long c = c0 + threadIdx.x;
    const bool un = (c < total) && !((covered[c >> 5] >> (c & 31)) & 1u);
    const unsigned bal = __ballot_sync(0xffffffffu, un);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    long long pre = base;
    for (int w = 0; w < warp; ++w) pre += wsum[w];
    if (un) {
      const long long pos = pre + __popc(bal & ((1u << lane) - 1u));
      if (soc) { soc[2 * pos] = (int)(c % n); soc[2 * pos + 1] = (int)(c / n); }
    }
    __syncthreads();
    if (threadIdx.x == 0) { long long t = 0; for (int w = 0; w < nw; ++w) t += wsum[w]; base += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *nsoc = base;
}

// generate_violated_Shor_minors (OMC.jl:2614-2640): score of every candidate minor, sum_t |Xt[i1,j1] Xt[i2,j2] - Xt[i1,j2] Xt[i2,j1]|,
// one thread per candidate; minors already in the node (sorted 64-bit keys, binary search) get -1.  Products and the difference
// are rounded separately (no FMA contraction), so the scores equal the reference's expression bit for bit and the top-n order
// does not depend on the device.  Xt: k column-major n x m slices.
__device__ __forceinline__ unsigned long long shor_key(int i1, int i2, int j1, int j2) {
  return ((unsigned long long)i1 << 48) | ((unsigned long long)i2 << 32) | ((unsigned long long)j1 << 16) | (unsigned long long)j2;
}
__global__ void __launch_bounds__(256) shor_score_kernel(const int* __restrict__ cand, long long ncand, const double* __restrict__ Xt, int k, int n, int m,
                                                         const unsigned long long* __restrict__ excl, long long nexcl, double* __restrict__ score) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= ncand) return;
  const int i1 = cand[4 * q], i2 = cand[4 * q + 1], j1 = cand[4 * q + 2], j2 = cand[4 * q + 3];
  const unsigned long long key = shor_key(i1, i2, j1, j2);
  long long lo = 0, hi = nexcl;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (excl[mid] < key) lo = mid + 1; else hi = mid;
  }
  if (lo < nexcl && excl[lo] == key) { score[q] = -1.0; return; }
  double sacc = 0.0;
  for (int t = 0; t < k; ++t) {
    const double* X = Xt + (size_t)t * n * m;
    const double x11 = X[(size_t)j1 * n + i1], x22 = X[(size_t)j2 * n + i2], x12 = X[(size_t)j2 * n + i1], x21 = X[(size_t)j1 * n + i2];
    sacc = __dadd_rn(sacc, fabs(__dsub_rn(__dmul_rn(x11, x22), __dmul_rn(x12, x21))));
  }
  score[q] = sacc;
}

}  // namespace omc
