/* oracle/oracle_bwt.c — TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C CPU restatement of what the reference's forward-BWT path computes, used ONLY as the
 * checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product path
 * (bwtc_b200/) never links, loads or calls anything in oracle/.
 *
 * Parity is PINNED: tests/test_oracle.py checks this file (a) against the known answers captured
 * from the compiled reference (committed in tests/golden/), and (b) byte-for-byte
 * against oracle/_ref/libbwtc_ref.so (the unmodified reference compiled here by oracle/Makefile),
 * both engines 'd' (divsufsort) and 's' (SA-IS), on seeded random inputs.
 *
 * What is restated (reference file:line):
 *   - the suffix order divsufsort/sais realise: all N suffixes of T[0..N), a suffix that is a proper
 *     prefix of another sorts first (implicit end-of-string sentinel)      divsufsort.c:38-192, sais.hxx:777-830
 *   - the output convention of the modified divbwt/divbwtf:
 *       pidx = rank of suffix 0; U[r] = T[SA[r]-1] for r != pidx; U[pidx] untouched; ++freqs[U[r]]
 *                                                                          divsufsort.c:440-522 (copy loop :506-512)
 *       n <= 1 early-out (U[0]=T[0], LFpowers untouched)                   divsufsort.c:488-489
 *       LFpowers[0] = pidx; LFpowers[j] = rank of suffix n - j*x, x = n / nLFpowers
 *                                                                          divsufsort.c:337-338,350,381,390,498-504
 *   - the block-level wrapper: reverse, save *end, *end = 0, raw transform over size+1 bytes,
 *     hole fill begin[LF[0]] = *end, restore *end                          bwtransforms/BWTransform.cpp:52-64
 *   - LFpowers sizing: 1 if length <= 256 (or starts == 0), else min(starts, 256)   BWTBlock.cpp:104-108
 *     with the manager clamping starts to [1,256]                          bwtransforms/BWTManager.cpp:60-64
 *
 * The suffix array is built by a textbook Manber-Myers prefix doubling with two counting-sort
 * passes per round — deliberately NOT the GPU algorithm's key packing, and not divsufsort.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Suffix array of T[0..N) under "shorter first".  SA and ISA must hold N entries.
 * Returns 0, or -2 on allocation failure. */
int oracle_suffix_array(const uint8_t* T, uint32_t N, uint32_t* SA, uint32_t* ISA) {
  if (N == 0) return 0;
  uint32_t* tmp = (uint32_t*)malloc((size_t)N * sizeof(uint32_t));
  uint32_t* nrk = (uint32_t*)malloc((size_t)N * sizeof(uint32_t));
  size_t nb = (size_t)N + 2 > 258 ? (size_t)N + 2 : 258;
  uint32_t* cnt = (uint32_t*)malloc(nb * sizeof(uint32_t));
  if (!tmp || !nrk || !cnt) { free(tmp); free(nrk); free(cnt); return -2; }

  /* round 0: order by first character; rank = index of the group head */
  memset(cnt, 0, 257 * sizeof(uint32_t));
  for (uint32_t i = 0; i < N; ++i) cnt[T[i] + 1]++;
  for (int c = 0; c < 256; ++c) cnt[c + 1] += cnt[c];
  {
    uint32_t head[256];
    for (int c = 0; c < 256; ++c) head[c] = cnt[c];
    for (uint32_t i = 0; i < N; ++i) { SA[cnt[T[i]]++] = i; }
    for (uint32_t i = 0; i < N; ++i) ISA[i] = head[T[i]];
  }

  for (uint32_t h = 1;; h <<= 1) {
    /* key2(i) = rank of suffix i+h shifted by one, 0 when i+h is past the end (the sentinel) */
    /* pass 1: counting sort of all i by key2 */
    memset(cnt, 0, ((size_t)N + 2) * sizeof(uint32_t));
    for (uint32_t i = 0; i < N; ++i) {
      uint32_t k2 = ((uint64_t)i + h < N) ? ISA[i + h] + 1 : 0;
      cnt[k2 + 1]++;
    }
    for (uint32_t k = 0; k <= N; ++k) cnt[k + 1] += cnt[k];
    for (uint32_t i = 0; i < N; ++i) {
      uint32_t k2 = ((uint64_t)i + h < N) ? ISA[i + h] + 1 : 0;
      tmp[cnt[k2]++] = i;
    }
    /* pass 2: stable counting sort by key1 = ISA[i] (a group-head index, so the bucket start) */
    memset(cnt, 0, ((size_t)N + 2) * sizeof(uint32_t));
    for (uint32_t i = 0; i < N; ++i) cnt[ISA[i] + 1]++;
    for (uint32_t k = 0; k <= N; ++k) cnt[k + 1] += cnt[k];
    for (uint32_t j = 0; j < N; ++j) { uint32_t i = tmp[j]; SA[cnt[ISA[i]]++] = i; }
    /* re-rank */
    int all_unique = 1;
    uint32_t head = 0;
    for (uint32_t j = 0; j < N; ++j) {
      if (j > 0) {
        uint32_t a = SA[j - 1], b = SA[j];
        uint32_t a2 = ((uint64_t)a + h < N) ? ISA[a + h] + 1 : 0;
        uint32_t b2 = ((uint64_t)b + h < N) ? ISA[b + h] + 1 : 0;
        if (ISA[a] != ISA[b] || a2 != b2) head = j; else all_unique = 0;
      }
      nrk[SA[j]] = head;
    }
    memcpy(ISA, nrk, (size_t)N * sizeof(uint32_t));
    if (all_unique || h >= N) break;
  }
  free(tmp); free(nrk); free(cnt);
  return 0;
}

/* The raw contract (Divsufsorter.hpp:54-65 -> divbwt/divbwtf, divsufsort.c:440-522).
 * T: N input bytes; U: N output bytes (may alias T); LF: nLF entries; freqs: 256 counters that are
 * INCREMENTED, or NULL.  Returns pidx (>= 0), N for N <= 1, -1 on bad arguments, -2 on alloc failure. */
int64_t oracle_bwt_raw(const uint8_t* T, uint8_t* U, uint32_t N, uint32_t* LF, uint32_t nLF,
                       uint32_t* freqs) {
  if (!T || !U) return -1;
  if (N <= 1) { if (N == 1) U[0] = T[0]; return N; }   /* divsufsort.c:488-489 */
  if (nLF == 0 || nLF > N) return -1;                  /* x = n/nLF would be 0: UB in the reference */
  uint32_t* SA = (uint32_t*)malloc((size_t)N * 4);
  uint32_t* ISA = (uint32_t*)malloc((size_t)N * 4);
  uint8_t* B = (uint8_t*)malloc(N);
  if (!SA || !ISA || !B) { free(SA); free(ISA); free(B); return -2; }
  if (oracle_suffix_array(T, N, SA, ISA)) { free(SA); free(ISA); free(B); return -2; }
  uint32_t pidx = ISA[0];
  for (uint32_t r = 0; r < N; ++r) B[r] = SA[r] ? T[SA[r] - 1] : 0;
  LF[0] = pidx;                                         /* divsufsort.c:500,503 */
  uint32_t x = N / nLF;                                 /* divsufsort.c:337 */
  for (uint32_t j = 1; j < nLF; ++j) LF[j] = ISA[N - j * x];  /* :350,381,390 */
  for (uint32_t r = 0; r < N; ++r) {                    /* divsufsort.c:506-512 */
    if (r != pidx) { U[r] = B[r]; if (freqs) ++freqs[B[r]]; }
  }
  free(SA); free(ISA); free(B);
  return pidx;
}

/* Number of starting points a block of length n gets (BWTManager.cpp:60-64 + BWTBlock.cpp:104-108). */
uint32_t oracle_num_starting_points(uint32_t n, uint32_t starts) {
  if (starts < 1) starts = 1; else if (starts > 256) starts = 256;
  return n <= 256 ? 1 : starts;
}

/* The block contract (BWTManager.cpp:53-58 -> BWTransform.cpp:52-64).  buf: n block bytes followed
 * by one writable slot whose value is preserved.  LF must hold 256 entries; *nLF receives the count.
 * freqs may be NULL. */
int64_t oracle_bwt_block(uint8_t* buf, uint32_t n, uint32_t starts, uint32_t* LF, uint32_t* nLF,
                         uint32_t* freqs) {
  if (!buf || n == 0) return -1;
  uint32_t k = oracle_num_starting_points(n, starts);
  *nLF = k;
  for (uint32_t a = 0, b = n - 1; a < b; ++a, --b) { uint8_t t = buf[a]; buf[a] = buf[b]; buf[b] = t; }
  uint8_t next = buf[n];
  buf[n] = 0;
  int64_t pidx = oracle_bwt_raw(buf, buf, n + 1, LF, k, freqs);
  if (pidx >= 0) buf[LF[0]] = buf[n];
  buf[n] = next;
  return pidx;
}

/* ---- inverse transform -----------------------------------------------------------------
 * Raw virtual InverseBWTransform::doTransform(byte* bwt, uint32 N, LFpow) restated as the plain LF walk of
 * FastInverseBWTransform (InverseBWT.cpp:58-115): count[] with the end-of-block row counted as the smallest symbol
 * (:74-77), rank of every row among equal characters (:79-91), prefix sums (:92), then follow
 * position -> count[ch] + rank from row 0 until the end-of-block row is reached, writing the original bytes in
 * forward order (:100-110).  MtlSaInverseBWTransform (what giveInverseTransformer returns, InverseBWT.cpp:42-45)
 * computes the same function with several chains; tests/test_oracle.py pins this restatement on it.
 * bwt: N bytes L[0..N) with bwt[eob] ignored; on return bwt[0..N-1) is the original block.  Returns N-1 or < 0. */
int64_t oracle_inverse_raw(uint8_t* bwt, uint32_t N, uint32_t eob) {
  if (!bwt || N < 1 || eob >= N) return -1;
  uint32_t* rank = (uint32_t*)malloc((size_t)N * 4);
  uint8_t* L = (uint8_t*)malloc(N);
  if (!rank || !L) { free(rank); free(L); return -2; }
  memcpy(L, bwt, N);
  uint64_t count[258];
  memset(count, 0, sizeof count);
  count[0] = 1;  /* the end-of-block symbol */
  rank[eob] = 0;
  for (uint32_t p = 0; p < N; ++p)
    if (p != eob) rank[p] = (uint32_t)count[(uint32_t)L[p] + 1]++;
  uint64_t acc = 0;
  for (int c = 0; c < 257; ++c) { uint64_t t = count[c]; count[c] = acc; acc += t; }  /* count[c+1] = symbols < c */
  uint32_t index = 0, position = 0;
  while (position != eob && index < N - 1) {
    uint8_t ch = L[position];
    bwt[index++] = ch;
    position = (uint32_t)(count[(uint32_t)ch + 1] + rank[position]);
  }
  free(rank); free(L);
  return index == N - 1 ? (int64_t)index : -3;  /* -3: not a valid transform (the cycle closed early) */
}

/* Block level: InverseBWTransform::doTransform(BWTBlock&) (InverseBWT.cpp:47-51) re-opens the hole
 * (*block.end() = data[LFpowers[0]]) and inverts n+1 rows.  buf: n bytes + one writable slot (preserved). */
int64_t oracle_inverse_block(uint8_t* buf, uint32_t n, uint32_t eob) {
  if (!buf || n == 0 || eob > n) return -1;
  uint8_t next = buf[n];
  buf[n] = buf[eob];
  int64_t rc = oracle_inverse_raw(buf, n + 1, eob);
  buf[n] = next;
  return rc;
}
