/* bwtc_b200/tools/gen_inputs.c — deterministic synthetic workload generators (fixed seeds) for the
 * four input families BASELINE.json names: order-2 Markov text, 4-letter DNA-like
 * sequences, highly repetitive mutated repeats, uniform random bytes.  Plain C, own PRNG
 * (splitmix64-seeded xoshiro256**), so the same bytes are produced here and on the GPU box.
 * Built by __graft_entry__.build() into bwtc_b200/libbwtc_gen.so; used by tests/ and bench.py. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix64(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void rng_seed(rng_t* r, uint64_t seed) { for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed); }
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t* r) {
  uint64_t* s = r->s;
  uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return result;
}
static inline double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static double rng_normal(rng_t* r) {
  double u1 = rng_unit(r), u2 = rng_unit(r);
  if (u1 < 1e-300) u1 = 1e-300;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
/* Marsaglia-Tsang for shape >= 1, boosted for shape < 1 */
static double rng_gamma(rng_t* r, double a) {
  if (a < 1.0) {
    double u = rng_unit(r);
    if (u < 1e-300) u = 1e-300;
    return rng_gamma(r, a + 1.0) * pow(u, 1.0 / a);
  }
  double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x = rng_normal(r), v = 1.0 + c * x;
    if (v <= 0) continue;
    v = v * v * v;
    double u = rng_unit(r);
    if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
    if (log(u > 1e-300 ? u : 1e-300) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
  }
}

/* uniform random bytes */
void bwtc_gen_random(uint8_t* out, uint64_t n, uint64_t seed) {
  rng_t r; rng_seed(&r, seed);
  uint64_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t v = rng_next(&r); memcpy(out + i, &v, 8); }
  if (i < n) { uint64_t v = rng_next(&r); memcpy(out + i, &v, n - i); }
}

/* i.i.d. uniform over {A,C,G,T} */
void bwtc_gen_dna(uint8_t* out, uint64_t n, uint64_t seed) {
  static const char sym[4] = {'A', 'C', 'G', 'T'};
  rng_t r; rng_seed(&r, seed);
  uint64_t i = 0;
  while (i < n) {
    uint64_t v = rng_next(&r);
    for (int k = 0; k < 32 && i < n; ++k, v >>= 2) out[i++] = (uint8_t)sym[v & 3];
  }
}

/* random `period`-byte seed tiled to n bytes, then round(n*mut_rate) positions overwritten with
 * uniform random bytes (config 3: period 4096, mut_rate 0.001) */
void bwtc_gen_repetitive(uint8_t* out, uint64_t n, uint64_t seed, uint32_t period, double mut_rate) {
  rng_t r; rng_seed(&r, seed);
  if (period == 0) period = 1;
  uint8_t* tile = (uint8_t*)malloc(period);
  for (uint32_t i = 0; i < period; ++i) tile[i] = (uint8_t)(rng_next(&r) >> 56);
  for (uint64_t i = 0; i < n; i += period) memcpy(out + i, tile, (n - i < period) ? (size_t)(n - i) : period);
  free(tile);
  uint64_t muts = (uint64_t)((double)n * mut_rate + 0.5);
  for (uint64_t m = 0; m < muts && n; ++m) {
    uint64_t v = rng_next(&r);
    out[(v >> 8) % n] = (uint8_t)(v & 0xFF);
  }
}

/* order-2 Markov text over `sigma` symbols mapped to bytes base..base+sigma-1; each of the sigma^2
 * contexts has transition probabilities ~ Dirichlet(alpha) (config 1/5: sigma 64, base 32, alpha 0.05).
 * `skip` symbols of the chain are generated and discarded first so that a long stream can be produced
 * piecewise only by regenerating from the start; callers that need block b of a stream instead use a
 * distinct seed per block (what bench.py does) — the table depends on `table_seed` only. */
void bwtc_gen_markov2(uint8_t* out, uint64_t n, uint64_t table_seed, uint64_t stream_seed,
                      uint32_t sigma, uint32_t base, double alpha) {
  if (sigma < 1) sigma = 1;
  if (sigma > 256) sigma = 256;
  rng_t r; rng_seed(&r, table_seed);
  size_t ctxs = (size_t)sigma * sigma;
  /* cumulative distribution per context as 32-bit thresholds */
  uint32_t* cdf = (uint32_t*)malloc(ctxs * sigma * sizeof(uint32_t));
  double* g = (double*)malloc(sigma * sizeof(double));
  for (size_t c = 0; c < ctxs; ++c) {
    double sum = 0;
    for (uint32_t s = 0; s < sigma; ++s) { g[s] = rng_gamma(&r, alpha); sum += g[s]; }
    if (sum <= 0) { g[0] = 1; sum = 1; }
    double acc = 0;
    for (uint32_t s = 0; s < sigma; ++s) {
      acc += g[s] / sum;
      double t = acc * 4294967296.0;
      cdf[c * sigma + s] = (t >= 4294967295.0) ? 0xFFFFFFFFu : (uint32_t)t;
    }
    cdf[c * sigma + sigma - 1] = 0xFFFFFFFFu;
  }
  free(g);
  rng_seed(&r, stream_seed ^ 0xA5A5A5A5DEADBEEFull);
  uint32_t a = (uint32_t)(rng_next(&r) % sigma), b = (uint32_t)(rng_next(&r) % sigma);
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t* row = cdf + ((size_t)a * sigma + b) * sigma;
    uint32_t u = (uint32_t)(rng_next(&r) >> 32);
    /* binary search for the first threshold >= u */
    uint32_t lo = 0, hi = sigma - 1;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (row[mid] >= u) hi = mid; else lo = mid + 1; }
    out[i] = (uint8_t)(base + lo);
    a = b; b = lo;
  }
  free(cdf);
}
