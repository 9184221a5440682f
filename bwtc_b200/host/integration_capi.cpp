/* bwtc_b200/host/integration_capi.cpp — extern "C" entry points over the integrated build (reference objects + patched
 * BWTManager / giveTransformer + bwtc::CudaBWTransform + bwtc::PipelinedCompressor) so that tests/ and bench.py can drive
 * it through ctypes.  The same calls a bwtc maintainer makes from compress.cpp:192-195. */
#define MAIN /* defines bwtc::verbosity in this TU (globaldefs.hpp:33-41) */
#include "globaldefs.hpp"

#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <vector>

#include "BWTBlock.hpp"
#include "Compressor.hpp"
#include "Decompressor.hpp"
#include "bwtransforms/BWTManager.hpp"
#include "bwtransforms/BWTransform.hpp"
#include "CudaBWTransform.hpp"
#include "MemStreams.hpp"
#include "PipelinedCompressor.hpp"
#include "RunStatistics.hpp"
#include "Utils.hpp"

namespace {
void put_err(char* err, unsigned errlen, const char* what) {
  if (err && errlen) { strncpy(err, what, errlen - 1); err[errlen - 1] = 0; }
}
}  // namespace

extern "C" {

/* BWTManager m; m.setStartingPoints(starts); m.initialize(choice); m.doTransform(block, freqs)  (BWTManager.cpp:53-58).
 * choice 'c' = CudaBWTransform through the patched manager; 'd' / 's' = the reference's CPU engines.  buf: n+1 bytes. */
int b200_manager_block(unsigned char* buf, unsigned n, unsigned starts, char choice, unsigned* LF_out, unsigned* nLF_out,
                       unsigned* freqs, char* err, unsigned errlen) {
  try {
    bwtc::verbosity = 0;
    if (!bwtc::BWTManager::isValidChoice(choice)) throw std::invalid_argument("invalid BWT choice");
    bwtc::BWTManager m;
    m.setStartingPoints(starts);
    m.initialize(choice);
    bwtc::BWTBlock b(buf, n, false);
    if (freqs) m.doTransform(b, freqs); else m.doTransform(b);
    if (!b.isTransformed()) throw std::runtime_error("block not flagged transformed");
    *nLF_out = (unsigned)b.LFpowers().size();
    for (size_t i = 0; i < b.LFpowers().size(); ++i) LF_out[i] = b.LFpowers()[i];
    return 0;
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}

/* giveTransformer(choice) + the NON-virtual base wrapper BWTransform::doTransform(BWTBlock&, freqs) (BWTransform.cpp:52-64):
 * host-side std::reverse / sentinel / hole fill around the raw virtual — INTEGRATION.md option B, what an unpatched
 * BWTManager would run.  buf: n+1 writable bytes. */
int b200_base_wrapper_block(unsigned char* buf, unsigned n, unsigned starts, char choice, unsigned* LF_out, unsigned* nLF_out,
                            unsigned* freqs, char* err, unsigned errlen) {
  try {
    bwtc::verbosity = 0;
    bwtc::BWTransform* t = bwtc::giveTransformer(choice);
    bwtc::BWTManager sizing;
    sizing.setStartingPoints(starts);
    bwtc::BWTBlock b(buf, n, false);
    b.prepareLFpowers(sizing.getStartingPoints());
    if (freqs) t->doTransform(b, freqs); else t->doTransform(b);
    delete t;
    *nLF_out = (unsigned)b.LFpowers().size();
    for (size_t i = 0; i < b.LFpowers().size(); ++i) LF_out[i] = b.LFpowers()[i];
    return 0;
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}

/* A manager / transformer that lives across blocks, as Compressor's m_bwtmanager does (Compressor.hpp:118): what the
 * per-block throughput of options B and C is measured with (the one-shot entry points above build and destroy the
 * CUDA context per call). */
void* b200_manager_new(char choice, unsigned starts) {
  try {
    bwtc::BWTManager* m = new bwtc::BWTManager();
    m->setStartingPoints(starts);
    m->initialize(choice);
    return m;
  } catch (...) { return 0; }
}
void b200_manager_free(void* h) { delete static_cast<bwtc::BWTManager*>(h); }
int b200_manager_transform(void* h, unsigned char* buf, unsigned n, unsigned* LF_out, unsigned* nLF_out, unsigned* freqs, char* err,
                           unsigned errlen) {
  try {
    bwtc::BWTBlock b(buf, n, false);
    static_cast<bwtc::BWTManager*>(h)->doTransform(b, freqs);
    *nLF_out = (unsigned)b.LFpowers().size();
    for (size_t i = 0; i < b.LFpowers().size(); ++i) LF_out[i] = b.LFpowers()[i];
    return 0;
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}
void* b200_transformer_new(char choice) {
  try { return bwtc::giveTransformer(choice); } catch (...) { return 0; }
}
void b200_transformer_free(void* h) { delete static_cast<bwtc::BWTransform*>(h); }
/* the NON-virtual base wrapper (host std::reverse, BWTransform.cpp:52-64) on a transformer that is kept */
int b200_transformer_base_wrapper(void* h, unsigned char* buf, unsigned n, unsigned starts, unsigned* LF_out, unsigned* nLF_out,
                                  unsigned* freqs, char* err, unsigned errlen) {
  try {
    bwtc::BWTManager sizing;
    sizing.setStartingPoints(starts);
    bwtc::BWTBlock b(buf, n, false);
    b.prepareLFpowers(sizing.getStartingPoints());
    static_cast<bwtc::BWTransform*>(h)->doTransform(b, freqs);
    *nLF_out = (unsigned)b.LFpowers().size();
    for (size_t i = 0; i < b.LFpowers().size(); ++i) LF_out[i] = b.LFpowers()[i];
    return 0;
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}

/* giveTransformer(choice)->doTransform(T, N, LF[, freqs]): the raw virtual on a caller-prepared buffer, as the
 * reference's tests call it (test/InverseBwtTest.cpp:57-66). */
int b200_transformer_raw(unsigned char* T, unsigned N, unsigned nLF, char choice, unsigned* LF_out, unsigned* freqs, char* err,
                         unsigned errlen) {
  try {
    bwtc::verbosity = 0;
    bwtc::BWTransform* t = bwtc::giveTransformer(choice);
    std::vector<bwtc::uint32> lf(nLF);
    if (freqs) t->doTransform(T, N, lf, freqs); else t->doTransform(T, N, lf);
    for (size_t i = 0; i < lf.size(); ++i) LF_out[i] = lf[i];
    delete t;
    return 0;
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}

/* CudaBWTransform::doTransformFused(vector<BWTBlock*>&, starts, freqs): the slices of one precompressor block in one call. */
int b200_fused_blocks(unsigned char** bufs, const unsigned* sizes, unsigned count, unsigned starts, unsigned* LF_out,
                      unsigned* nLF_out, unsigned* freqs, char* err, unsigned errlen) {
  try {
    bwtc::verbosity = 0;
    std::vector<bwtc::BWTBlock> store;
    store.reserve(count);
    for (unsigned k = 0; k < count; ++k) store.push_back(bwtc::BWTBlock(bufs[k], sizes[k], false));
    std::vector<bwtc::BWTBlock*> blocks;
    for (unsigned k = 0; k < count; ++k) blocks.push_back(&store[k]);
    bwtc::CudaBWTransform t;
    t.doTransformFused(blocks, starts, reinterpret_cast<bwtc::uint32 (*)[256]>(freqs));
    for (unsigned k = 0; k < count; ++k) {
      if (!store[k].isTransformed()) throw std::runtime_error("block not flagged transformed");
      nLF_out[k] = (unsigned)store[k].LFpowers().size();
      for (size_t i = 0; i < store[k].LFpowers().size(); ++i) LF_out[k * 256 + i] = store[k].LFpowers()[i];
    }
    return 0;
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}

/* The reference's own synchronous Compressor (Compressor.cpp:65-120) with any BWT choice of the patched manager. */
long long b200_sync_compress_file(const char* in, const char* out, unsigned long long memLimit, char coder, char choice,
                                  unsigned starts, const char* prepr, char* err, unsigned errlen) {
  try {
    bwtc::verbosity = 0;
    size_t sz;
    {
      bwtc::Compressor c(std::string(in), std::string(out), std::string(prepr ? prepr : ""), (size_t)memLimit, coder);
      c.initializeBwtAlgorithm(choice, starts);
      sz = c.compress(1);
    }
    return (long long)sz;
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}

/* PipelinedCompressor on files.  timings: 10 doubles (total, reader_busy, encoder_busy_sum, writer_busy, bwt_wait_sum,
 * precompressorBlocks, bwtBlocks, inputBytes, encoderThreads, 0) or NULL. */
long long b200_pipelined_compress_file(const char* in, const char* out, unsigned long long memLimit, char coder, char choice,
                                       unsigned starts, const char* prepr, unsigned threads, unsigned lookahead,
                                       const int* devices, unsigned ndevices, int depth, unsigned rank, unsigned world,
                                       double* timings, char* err, unsigned errlen) {
  try {
    bwtc::verbosity = 0;
    size_t sz;
    {
      bwtc::PipelinedCompressor c(std::string(in), std::string(out), std::string(prepr ? prepr : ""), (size_t)memLimit, coder);
      c.initializeBwtAlgorithm(choice, starts);
      c.setDevices(std::vector<int>(devices, devices + ndevices), depth);
      c.setLookahead(lookahead);
      c.setShard(rank, world);
      sz = c.compress(threads);
      if (timings) {
        const bwtc::PipelineTimings& t = c.timings();
        double v[10] = {t.total, t.reader_busy, t.encoder_busy_sum, t.writer_busy, t.bwt_wait_sum, (double)t.precompressorBlocks,
                        (double)t.bwtBlocks, (double)t.inputBytes, (double)t.encoderThreads, 0};
        memcpy(timings, v, sizeof v);
      }
    }
    return (long long)sz;
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}

/* PipelinedCompressor from caller memory (bench: no file system in the timed region).  out_buf == NULL: the compressed
 * bytes are only counted; otherwise they are copied to out_buf (capacity out_cap; -2 if it does not fit). */
long long b200_pipelined_compress_mem(const unsigned char* data, unsigned long long size, unsigned long long memLimit, char coder,
                                      char choice, unsigned starts, unsigned threads, unsigned lookahead, const int* devices,
                                      unsigned ndevices, int depth, unsigned char* out_buf, unsigned long long out_cap,
                                      double* timings, char* err, unsigned errlen) {
  try {
    bwtc::verbosity = 0;
    bwtc::MemOutStream* mem = out_buf ? new bwtc::MemOutStream() : 0;
    bwtc::CountingOutStream* cnt = out_buf ? 0 : new bwtc::CountingOutStream();
    bwtc::OutStream* os = out_buf ? static_cast<bwtc::OutStream*>(mem) : static_cast<bwtc::OutStream*>(cnt);
    bwtc::PipelinedCompressor c(new bwtc::MemInStream(data, (size_t)size), os, std::string(""), (size_t)memLimit, coder);
    c.initializeBwtAlgorithm(choice, starts);
    c.setDevices(std::vector<int>(devices, devices + ndevices), depth);
    c.setLookahead(lookahead);
    const size_t sz = c.compress(threads);
    if (timings) {
      const bwtc::PipelineTimings& t = c.timings();
      double v[10] = {t.total, t.reader_busy, t.encoder_busy_sum, t.writer_busy, t.bwt_wait_sum, (double)t.precompressorBlocks,
                      (double)t.bwtBlocks, (double)t.inputBytes, (double)t.encoderThreads, 0};
      memcpy(timings, v, sizeof v);
    }
    if (out_buf) {
      if (mem->data().size() > out_cap) return -2;
      if (!mem->data().empty()) memcpy(out_buf, &mem->data()[0], mem->data().size());
    }
    return (long long)sz;  /* the streams are deleted by the compressor's destructor */
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}

long long b200_merge_parts(const char* const* parts, unsigned nparts, const char* out, char coder, char* err, unsigned errlen) {
  try {
    std::vector<std::string> p;
    for (unsigned i = 0; i < nparts; ++i) p.push_back(parts[i]);
    return (long long)bwtc::PipelinedCompressor::mergeParts(p, out, coder);
  } catch (const std::exception& e) { put_err(err, errlen, e.what()); return -1; }
}

long long b200_uncompress_file(const char* in, const char* out) {
  bwtc::verbosity = 0;
  size_t sz;
  {
    bwtc::Decompressor d((std::string(in)), (std::string(out)));
    sz = d.decompress(1);
  }
  return (long long)sz;
}

/* Host side of the GPU run statistics (RunStatistics.cpp), checked without a GPU: the maximal runs of data[0..n) are
 * computed here the plain way and published as the engine would publish them; then every section is asked for through
 * utils::calculateRunFrequenciesAndStoreRuns (wrapped at link time -> answered from the registry) and through the
 * reference's real implementation.  Returns 0 if all sections agree (freqs, run symbols, run lengths, counts), the
 * 1-based index of the first differing section otherwise, -1 if the registry did not answer. */
extern "C" bwtc::uint64 __real__ZN5utils35calculateRunFrequenciesAndStoreRunsEPmPhPjPKhm(bwtc::uint64*, bwtc::byte*, bwtc::uint32*,
                                                                                       const bwtc::byte*, size_t);
int b200_test_run_slicing(const unsigned char* data, unsigned n, const unsigned* section_len, unsigned nsections) {
  bwtc::runstats::Record* rec = new bwtc::runstats::Record();
  rec->begin = data;
  rec->size = n;
  for (unsigned i = 0; i < n; ++i)
    if (i == 0 || data[i] != data[i - 1]) { rec->symbol.push_back(data[i]); rec->start.push_back(i); }
  bwtc::runstats::publish(rec);
  const size_t served0 = bwtc::runstats::served();
  std::vector<bwtc::byte> sa(n + 1), sb(n + 1);
  std::vector<bwtc::uint32> la(n + 1), lb(n + 1);
  unsigned beg = 0;
  int bad = 0;
  for (unsigned s = 0; s < nsections && !bad; ++s) {
    const unsigned len = section_len[s];
    if (len == 0) continue;
    bwtc::uint64 fa[256], fb[256];
    memset(fa, 0, sizeof fa);
    memset(fb, 0, sizeof fb);
    const bwtc::uint64 na = utils::calculateRunFrequenciesAndStoreRuns(fa, &sa[0], &la[0], data + beg, len);
    const bwtc::uint64 nb = __real__ZN5utils35calculateRunFrequenciesAndStoreRunsEPmPhPjPKhm(fb, &sb[0], &lb[0], data + beg, len);
    if (na != nb || memcmp(fa, fb, sizeof fa) != 0 || memcmp(&sa[0], &sb[0], na) != 0 || memcmp(&la[0], &lb[0], na * 4) != 0) bad = (int)s + 1;
    beg += len;
  }
  const bool answered = bwtc::runstats::served() > served0;
  bwtc::runstats::release(data);
  if (bad) return bad;
  return answered ? 0 : -1;
}

size_t b200_run_statistics_served(void) { return bwtc::runstats::served(); }

/* Releases the look-ahead pipelines (device scratch, worker threads) kept between compress() calls. */
void b200_shutdown(void) { bwtc::CudaBWTransform::shutdownLookahead(); }

int b200_is_valid_choice(char c) { return bwtc::BWTManager::isValidChoice(c) ? 1 : 0; }

} /* extern "C" */
