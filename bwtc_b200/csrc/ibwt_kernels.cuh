// bwtc_b200/csrc/ibwt_kernels.cuh — hand-written sm_100a kernels of the INVERSE Burrows-Wheeler transform
// (SURVEY.md §8f row f4).  Replaces InverseBWTransform::doTransform / MtlSaInverseBWTransform
// (bwtransforms/InverseBWT.hpp:45-55, InverseBWT.cpp:47-51, MtlSaInverseBWT.cpp:246-362) — not a port of MTL-SA.
//
// Convention (same as the forward path): T' = reverse(X) . $, N = n + 1, L[r] = T'[SA[r] - 1],
// eob = LFpowers[0] = rank of suffix 0 (L[eob] is the sentinel).  The reference inverts by following LF from row 0
// (FastInverseBWTransform, InverseBWT.cpp:58-115) — one dependent random access per byte, inherently serial; its MTL-SA
// variant runs a handful of such chains (the LFpowers starting points) on CPU threads.  A GPU needs ~10^5 independent
// chains, so the starting points are made on the device instead:
//   1. k_inv_keys     key[r] = L[r] + 1 (0 for the sentinel row), 9 significant bits, + both digit histograms
//   2. k_radix_pass   x 2 (the forward path's one-sweep LSD pass, stable): Psi[g] = the L-row of F-row g, i.e. the
//                     rank of the suffix that starts one position LATER in T' — walking Psi walks T' forwards
//   3. k_inv_ctable   C[c] = number of key values < c (F[g] is recovered by a search in C, no F array in memory)
//   4. k_inv_walk1    every K-th rank is a SPLITTER; each splitter follows Psi to the next splitter: (next, length)
//   5. k_inv_jump     x ceil(log2 S): pointer jumping over the S = N/K splitters -> distance of each to the end of the
//                     cycle opened at rank 0 (suffix N-1), i.e. its text position
//   6. k_inv_walk2    each splitter walks its stretch again and writes X[n-1-q] = F[rank(q)] for its text positions q
// Two passes of N dependent-per-chain but mutually independent random reads with ~N/K chains in flight: bound by the
// L2/DRAM random-access rate, not by latency.  LFpowers[1..] are not needed (the device makes far denser samples).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bwtc_b200 {

constexpr uint32_t INV_K = 128;          // splitter spacing in rank space (average chain length)
constexpr uint32_t INV_NIL = 0xFFFFFFFFu;

// key[r] = L[r] + 1, 0 at the sentinel row.  Row N-1 holds, in the block contract, the byte the forward transform
// moved into the hole (out[eob] = L[N-1], BWTransform.cpp:60); in the raw contract in[N-1] itself.
// hist[0..255] = histogram of key & 0xFF, hist[256..511] = histogram of key >> 8.
__global__ void __launch_bounds__(256) k_inv_keys(const uint8_t* __restrict__ in, uint32_t N, uint32_t eob, int block_mode,
                                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_lo[256];
  __shared__ uint32_t s_hi;
  s_lo[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_hi = 0;
  __syncthreads();
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
    uint32_t key;
    if (r == eob) key = 0;
    else if (r == N - 1u && block_mode) key = (uint32_t)in[eob] + 1u;
    else key = (uint32_t)in[r] + 1u;
    keys[r] = key;
    atomicAdd(&s_lo[key & 0xFFu], 1u);
    if (key >> 8) atomicAdd(&s_hi, 1u);
  }
  __syncthreads();
  const uint32_t v = s_lo[threadIdx.x];
  if (v) atomicAdd(&hist[threadIdx.x], v);
  if (threadIdx.x == 0) {
    if (s_hi) atomicAdd(&hist[256 + 1], s_hi);
  }
}

// hist[256 + 0] = N - hist[256 + 1] (the second digit pass needs the full histogram of its digit); C[0..257]:
// C[v] = number of keys < v, v = 0..257 (key values are 0..256).
__global__ void k_inv_ctable(uint32_t* __restrict__ hist, uint32_t N, uint32_t* __restrict__ C) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint32_t n256 = hist[256 + 1];  // keys with value 256 (byte 0xFF)
  hist[256 + 0] = N - n256;
  uint32_t acc = 0;
  for (uint32_t v = 0; v <= 256; ++v) {
    C[v] = acc;
    uint32_t cnt;
    if (v == 0) cnt = hist[0] - n256;      // low digit 0 is shared by key 0 (the sentinel) and key 256
    else if (v == 256) cnt = n256;
    else cnt = hist[v];
    acc += cnt;
  }
  C[257] = acc;
}

// Psi[g] = iota_top - vals[g] (the radix pass numbered the rows downwards, see k_radix_pass IOTA).
__device__ __forceinline__ uint32_t inv_psi(const uint32_t* __restrict__ vals, uint32_t iota_top, uint32_t g) {
  return iota_top - vals[g];
}

// Splitter s = rank s*K.  Follow Psi until the next splitter: nxt[s] = its index, len[s] = steps taken (>= 1).
// The splitter whose successor is splitter 0 gets nxt = INV_NIL: the cycle is opened at rank 0 (= suffix N-1).
__global__ void __launch_bounds__(256) k_inv_walk1(const uint32_t* __restrict__ vals, uint32_t N, uint32_t S,
                                                   uint32_t* __restrict__ nxt, uint32_t* __restrict__ len,
                                                   const uint32_t* __restrict__ ctrl) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S || ctrl[0]) return;  // ctrl[CTR_ERR]: a digit pass failed, Psi is not trustworthy (the host repeats the call)
  uint32_t g = s * INV_K, t = 0;
  do {
    g = inv_psi(vals, N - 1u, g);
    ++t;
  } while (g % INV_K != 0u && t < N);
  const uint32_t to = g / INV_K;
  nxt[s] = (to == 0u) ? INV_NIL : to;
  len[s] = t;
}

// One pointer-jumping round: dist_out[s] = dist_in[s] + dist_in[nxt_in[s]], nxt_out[s] = nxt_in[nxt_in[s]].
__global__ void __launch_bounds__(256) k_inv_jump(const uint32_t* __restrict__ nxt_in, const uint32_t* __restrict__ dist_in,
                                                  uint32_t S, uint32_t* __restrict__ nxt_out, uint32_t* __restrict__ dist_out) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const uint32_t nx = nxt_in[s];
  uint32_t d = dist_in[s], n2 = INV_NIL;
  if (nx != INV_NIL) {
    d += dist_in[nx];
    n2 = nxt_in[nx];
  }
  nxt_out[s] = n2;
  dist_out[s] = d;
}

// dist[s] = D(s) = steps from splitter s to the end of the opened cycle (back at rank 0), so the text position of
// splitter s is q0 = (2N - 1 - D(s)) mod N (splitter 0 = rank 0 = suffix N-1).  Walk len[s] steps: the row at step t is
// the suffix at text position q = q0 + t; its first character F[g] is X[n-1-q] (q = N-1 is the sentinel: nothing).
__global__ void __launch_bounds__(256) k_inv_walk2(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ len,
                                                   const uint32_t* __restrict__ dist, const uint32_t* __restrict__ C,
                                                   uint32_t N, uint32_t S, uint8_t* __restrict__ out,
                                                   const uint32_t* __restrict__ ctrl) {
  __shared__ uint32_t s_C[258];
  if (ctrl[0]) return;  // (the caller's buffer — possibly the in-place input — stays untouched)
  for (int i = threadIdx.x; i < 258; i += blockDim.x) s_C[i] = C[i];
  __syncthreads();
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const uint32_t n = N - 1u;
  uint32_t g = s * INV_K;
  const uint32_t L = len[s];
  uint32_t q = (uint32_t)((2ull * N - 1ull - (unsigned long long)dist[s]) % N);
  for (uint32_t t = 0; t < L; ++t) {
    if (q < n) {
      // F[g]: the largest key value v with C[v] <= g (binary search over the 257 bucket starts)
      uint32_t lo = 0, hi = 256;
      while (lo < hi) {
        const uint32_t mid = (lo + hi + 1u) >> 1;
        if (s_C[mid] <= g) lo = mid; else hi = mid - 1u;
      }
      out[n - 1u - q] = (uint8_t)(lo - 1u);
    }
    g = inv_psi(vals, N - 1u, g);
    q = (q + 1u == N) ? 0u : q + 1u;
  }
}

}  // namespace bwtc_b200
