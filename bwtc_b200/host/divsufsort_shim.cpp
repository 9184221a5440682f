/* bwtc_b200/host/divsufsort_shim.cpp — link-time drop-in for the reference's C entry points.
 *
 * Defines divbwt / divbwtf with EXACTLY the reference's signatures (bwtransforms/divsufsort.h:86-94) on top
 * of the C-ABI, so an unmodified bwtc build that links this object INSTEAD of bwtransforms/divsufsort.c,
 * sssort.c and trsort.c runs its Divsufsorter path (bwtransforms/Divsufsorter.hpp:54-65) on the GPU with no
 * source change at all (option A).  oracle/Makefile.cudaref builds exactly that as a parity
 * check: the resulting .bwtc files must be byte-identical to the reference's.
 *
 * A = scratch suffix array of the reference; ignored (the engine owns its scratch).  A failure cannot be
 * reported through Divsufsorter (it drops the return value), so it aborts loudly: no CPU fallback.
 */
#include <cstdio>
#include <cstdlib>

#include "../../include/bwtc_cuda.h"

namespace {
/* One context per calling thread (created on first use, regrown for larger blocks, released when the thread ends): callers
 * on different threads — e.g. the encoder threads of a parallel compressor — do not serialise on each other. */
struct ThreadCtx {
  bwtc_cuda_ctx* ctx;
  uint32_t cap;
  ThreadCtx() : ctx(0), cap(0) {}
  ~ThreadCtx() { if (ctx) bwtc_cuda_ctx_destroy(ctx); }
};
thread_local ThreadCtx t_ctx;

bwtc_cuda_ctx* ctx_for(uint32_t n) {
  ThreadCtx& t = t_ctx;
  if (t.ctx && n <= t.cap) return t.ctx;
  uint64_t want = n;
  if (t.ctx) {
    want = (uint64_t)t.cap * 2 > want ? (uint64_t)t.cap * 2 : want;
    bwtc_cuda_ctx_destroy(t.ctx);
    t.ctx = 0;
  }
  if (want < (1u << 20)) want = 1u << 20;
  if (want > BWTC_CUDA_MAX_BLOCK) want = BWTC_CUDA_MAX_BLOCK;
  const char* dev = getenv("BWTC_CUDA_DEVICE");
  int rc = bwtc_cuda_ctx_create(&t.ctx, dev ? atoi(dev) : 0, (uint32_t)want);
  if (rc != 0) {
    fprintf(stderr, "bwtc_b200 divsufsort shim: cannot create CUDA context (%d): %s\n", rc, bwtc_cuda_global_error());
    abort();
  }
  t.cap = (uint32_t)want;
  return t.ctx;
}
}  // namespace

extern "C" {

/* saidx_t is int32_t, sauchar_t is uint8_t (divsufsort.h:41-60) */
int32_t divbwtf(const uint8_t* T, uint8_t* U, int32_t* A, int32_t n, unsigned* LFpowers, unsigned nLFpowers,
                unsigned freqs[256]) {
  (void)A;
  if (T == 0 || U == 0 || n < 0) return -1;           /* divsufsort.c:488 */
  if (n <= 1) { if (n == 1) U[0] = T[0]; return n; }  /* divsufsort.c:489 */
  bwtc_cuda_ctx* c = ctx_for((uint32_t)n);
  int64_t rc = bwtc_cuda_divbwtf(c, T, U, (uint32_t)n, LFpowers, nLFpowers, freqs);
  if (rc < 0) {
    fprintf(stderr, "bwtc_b200 divsufsort shim: GPU transform failed (%lld): %s\n", (long long)rc, bwtc_cuda_last_error(c));
    abort();
  }
  return (int32_t)rc;
}

int32_t divbwt(const uint8_t* T, uint8_t* U, int32_t* A, int32_t n, unsigned* LFpowers, unsigned nLFpowers) {
  return divbwtf(T, U, A, n, LFpowers, nLFpowers, 0);
}

const char* divsufsort_version(void) { return "bwtc_b200 CUDA drop-in for libdivsufsort 2.0.0 (bwtc-modified)"; }

}  // extern "C"
