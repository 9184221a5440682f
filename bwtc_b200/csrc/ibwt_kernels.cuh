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
//   4. k_inv_walk1    every K-th rank (K = 32) is a SPLITTER; each splitter follows Psi to the next splitter: (next, length)
//   5. k_inv_jump     x ceil(log2 S): pointer jumping over the S = N/K splitters -> distance of each to the end of the
//                     cycle opened at rank 0 (suffix N-1), i.e. its text position
//   6. k_inv_walk2    each splitter walks its stretch again and writes X[n-1-q] = F[rank(q)] for its text positions q
// Two passes of N dependent-per-chain but mutually independent random reads with ~N/K chains in flight: bound by the
// L2/DRAM random-access rate, not by latency.  LFpowers[1..] are not needed (the device makes far denser samples).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bwtc_b200 {

#ifndef BWTC_INV_K
#define BWTC_INV_K 32
#endif
constexpr uint32_t INV_K = BWTC_INV_K;   // splitter spacing in rank space = average chain length.  Measured (32 MiB Markov, one
                                         // block): 128 -> 2.63 ms, 64 -> 2.29 ms, 32 -> 2.00 ms: shorter chains even out the
                                         // geometric spread of chain lengths that leaves lanes idle (profiles/r02_summary.md)
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

// =====================================================================================================
// Run statistics of the transformed block (the reference's own TODO at HuffmanCoders.cpp:54:
// "Also gather information about the runs during BWT").  HuffmanEncoder::encodeData re-scans every section of the BWT
// byte by byte to split it into runs (utils::calculateRunFrequenciesAndStoreRuns, Utils.cpp:150-170).  The device has
// the bytes in HBM anyway: two streaming kernels emit the maximal runs of the whole block — (symbol, start position)
// pairs in order — and the host side only slices them at section boundaries.  They are shipped to the host only when
// they are few (a repetitive block: 5 bytes per run instead of a scan over n bytes); for text-like blocks, where three
// out of four positions start a run, the scan on the host stays cheaper than the PCIe traffic.
//   k_run_count : heads per 4096-byte tile (head(i) = i == 0 || out[i] != out[i-1])
//   (k_scan_tile_counts, the forward path's scan, turns them into exclusive offsets)
//   k_run_emit  : symbol[off + k] = out[i], start[off + k] = i for the k-th head of the tile
// =====================================================================================================
constexpr uint32_t RUN_TILE = 4096;  // bytes per CTA (16 per thread)

__device__ __forceinline__ uint32_t run_heads16(const uint8_t* __restrict__ out, uint32_t n, uint32_t i0, uint8_t* b) {
  // 16 consecutive bytes of this thread + the byte before them; returns the head mask
  uint8_t prev = (i0 > 0 && i0 - 1 < n) ? out[i0 - 1] : 0;
  uint32_t mask = 0;
  if (i0 + 16 <= n && (reinterpret_cast<uintptr_t>(out + i0) & 15u) == 0) {
    const uint4 v = *reinterpret_cast<const uint4*>(out + i0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) b[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) b[k] = (i0 + k < n) ? out[i0 + k] : 0;
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t i = i0 + k;
    if (i < n && (i == 0 || b[k] != prev)) mask |= 1u << k;
    prev = b[k];
  }
  return mask;
}

__global__ void __launch_bounds__(256) k_run_count(const uint8_t* __restrict__ out, uint32_t n, uint32_t* __restrict__ tile_cnt,
                                                   const LadderState* __restrict__ st) {
  __shared__ uint32_t s_w[8];
  if (st->m != 0u) return;  // enqueued speculatively behind the ladder: the block is not finished yet
  uint8_t b[16];
  const uint32_t i0 = blockIdx.x * RUN_TILE + threadIdx.x * 16u;
  uint32_t c = __popc(run_heads16(out, n, i0, b));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    tile_cnt[blockIdx.x] = t;
  }
}

// total[0] = number of runs; the arrays are written only while the index stays below `capacity`.
__global__ void __launch_bounds__(256) k_run_emit(const uint8_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ tile_cnt,
                                                  const uint32_t* __restrict__ tile_excl, uint32_t ntiles, uint32_t capacity,
                                                  uint8_t* __restrict__ symbol, uint32_t* __restrict__ start,
                                                  uint32_t* __restrict__ total, const LadderState* __restrict__ st) {
  __shared__ uint32_t s_w[8];
  if (st->m != 0u) return;
  uint8_t b[16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i0 = blockIdx.x * RUN_TILE + threadIdx.x * 16u;
  const uint32_t mask = run_heads16(out, n, i0, b);
  const uint32_t c = __popc(mask);
  uint32_t inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_w[w];
  uint32_t pos = tile_excl[blockIdx.x] + woff + inc - c;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if ((mask >> k) & 1u) {
      if (pos < capacity) { symbol[pos] = b[k]; start[pos] = i0 + k; }
      ++pos;
    }
  }
  if (blockIdx.x == ntiles - 1u && threadIdx.x == 0) total[0] = tile_excl[blockIdx.x] + tile_cnt[blockIdx.x];
}

}  // namespace bwtc_b200
