/* oracle/ref_driver.cpp — TEST INFRASTRUCTURE, not product code.
 *
 * Thin extern "C" glue over the UNMODIFIED reference (pjmikkol/bwtc) sources, compiled
 * where they lie under $BWTC_REF (/root/reference) by oracle/Makefile into
 * oracle/_ref/libbwtc_ref.so.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product path never does.
 *
 * Entry points drive the reference exactly as its own callers do:
 *   ref_bwt_block    -> BWTManager::doTransform(BWTBlock&, uint32*)   bwtransforms/BWTManager.cpp:53-58
 *   ref_bwt_raw      -> BWTransform::doTransform(byte*,uint32,vector&[,freqs]) (the raw virtual,
 *                       Divsufsorter.hpp:54-65 / SA-IS-bwt.cpp:41-54), as test/InverseBwtTest.cpp:57-66 does
 *   ref_inverse_block-> InverseBWTransform::doTransform(BWTBlock&)     bwtransforms/InverseBWT.cpp:47-51
 *   ref_compress     -> Compressor(in,out,prepr,memLimit,coder)+initializeBwtAlgorithm+compress(1)
 *                       (Compressor.hpp:99-109, compress.cpp:192-195)
 *   ref_uncompress   -> Decompressor(in,out).decompress(1)             (Decompressor.cpp:58-94)
 */
#define MAIN /* defines bwtc::verbosity in this TU (globaldefs.hpp:33-41) */
#include "globaldefs.hpp"
#include "BWTBlock.hpp"
#include "Compressor.hpp"
#include "Decompressor.hpp"
#include "bwtransforms/BWTManager.hpp"
#include "bwtransforms/BWTransform.hpp"
#include "bwtransforms/InverseBWT.hpp"

#include <cstring>
#include <string>
#include <vector>

extern "C" {

/* buf must have n+1 writable bytes (PrecompressorBlock.cpp:41-49 allocates size+1).
 * LF_out must hold 256 entries.  freqs (256) may be NULL -> the no-freqs overload. */
int ref_bwt_block(unsigned char* buf, unsigned n, unsigned starts, char algo,
                  unsigned* LF_out, unsigned* nLF_out, unsigned* freqs) {
  bwtc::verbosity = 0;
  bwtc::BWTManager m;
  m.setStartingPoints(starts);
  m.initialize(algo);
  bwtc::BWTBlock b(buf, n, false);
  if (freqs) m.doTransform(b, freqs); else m.doTransform(b);
  std::vector<bwtc::uint32>& lf = b.LFpowers();
  *nLF_out = (unsigned)lf.size();
  for (size_t i = 0; i < lf.size(); ++i) LF_out[i] = lf[i];
  return 0;
}

/* T holds N bytes (caller-prepared reverse(text)+'\0' in the reference's tests, but any bytes are
 * legal).  nLF = LFpowers.size() chosen by the caller. */
int ref_bwt_raw(unsigned char* T, unsigned N, unsigned nLF, char algo,
                unsigned* LF_out, unsigned* freqs) {
  bwtc::verbosity = 0;
  bwtc::BWTransform* t = bwtc::giveTransformer(algo);
  std::vector<bwtc::uint32> lf(nLF);
  if (freqs) t->doTransform(T, N, lf, freqs); else t->doTransform(T, N, lf);
  for (size_t i = 0; i < lf.size(); ++i) LF_out[i] = lf[i];
  delete t;
  return 0;
}

/* buf: n transformed bytes (+1 writable slot), LF: starting points as written by the forward
 * transform.  Restores the original block in place. */
int ref_inverse_block(unsigned char* buf, unsigned n, const unsigned* LF, unsigned nLF) {
  bwtc::verbosity = 0;
  bwtc::BWTBlock b(buf, n, true);
  b.LFpowers().assign(LF, LF + nLF);
  bwtc::InverseBWTransform* inv = bwtc::giveInverseTransformer();
  inv->doTransform(b);
  delete inv;
  return 0;
}

long long ref_compress(const char* in, const char* out, unsigned long long memLimit,
                       char coder, char algo, unsigned starts) {
  bwtc::verbosity = 0;
  size_t sz;
  {
    bwtc::Compressor c(std::string(in), std::string(out), std::string(""), (size_t)memLimit, coder);
    c.initializeBwtAlgorithm(algo, starts);
    sz = c.compress(1);
  } /* streams are closed by the destructor (Compressor.cpp:49-53) */
  return (long long)sz;
}

long long ref_uncompress(const char* in, const char* out) {
  bwtc::verbosity = 0;
  size_t sz;
  {
    bwtc::Decompressor d((std::string(in)), (std::string(out)));
    sz = d.decompress(1);
  }
  return (long long)sz;
}

} /* extern "C" */
