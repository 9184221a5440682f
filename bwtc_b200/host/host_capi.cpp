/* bwtc_b200/host/host_capi.cpp — tiny C shim so tests/ can drive the C++ mirror classes through ctypes. */
#include <cstring>
#include <exception>
#include <stdexcept>
#include <vector>

#include "BWTransform.hpp"

extern "C" {

/* BWTManager m(starts); m.initialize('c'); BWTBlock b(buf, n, false); m.doTransform(b, freqs);
 * exactly the driver shape of the reference.  buf needs n+1 writable bytes only when
 * `via_base_wrapper` != 0 (then the REFERENCE's host-side convention BWTransform.cpp:52-64 runs around the raw
 * virtual).  Returns 0, or -1 with the exception text in err. */
int bwtc_host_manager_transform(unsigned char* buf, unsigned n, unsigned starts, int via_base_wrapper,
                                unsigned* LF_out, unsigned* nLF_out, unsigned* freqs, char* err, unsigned errlen) {
  try {
    bwtc_b200::BWTBlock b(buf, n, false);
    if (via_base_wrapper) {
      bwtc_b200::CudaBWTransform t;
      bwtc_b200::BWTManager sizing(1);
      sizing.setStartingPoints(starts);
      b.prepareLFpowers(sizing.getStartingPoints());
      if (freqs) t.bwtc_b200::BWTransform::doTransform(b, freqs);
      else t.bwtc_b200::BWTransform::doTransform(b);
    } else {
      bwtc_b200::BWTManager m(1);
      m.setStartingPoints(starts);
      m.initialize('c');
      if (freqs) m.doTransform(b, freqs); else m.doTransform(b);
    }
    if (!b.isTransformed()) throw std::runtime_error("block not flagged transformed");
    *nLF_out = (unsigned)b.LFpowers().size();
    for (size_t i = 0; i < b.LFpowers().size(); ++i) LF_out[i] = b.LFpowers()[i];
    return 0;
  } catch (const std::exception& e) {
    if (err && errlen) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return -1;
  }
}

/* BWTManager m(starts); m.initialize('c'); m.doTransform(blocks, freqs) — the batched extension: count blocks,
 * bufs[k] / sizes[k]; LF_out = count x 256, nLF_out = count, freqs = count x 256 or NULL. */
int bwtc_host_manager_transform_batch(unsigned char** bufs, const unsigned* sizes, unsigned count, unsigned starts,
                                      unsigned* LF_out, unsigned* nLF_out, unsigned* freqs, char* err, unsigned errlen) {
  try {
    std::vector<bwtc_b200::BWTBlock> store;
    store.reserve(count);
    std::vector<bwtc_b200::BWTBlock*> blocks;
    for (unsigned k = 0; k < count; ++k) { store.push_back(bwtc_b200::BWTBlock(bufs[k], sizes[k], false)); }
    for (unsigned k = 0; k < count; ++k) blocks.push_back(&store[k]);
    bwtc_b200::BWTManager m(1);
    m.setStartingPoints(starts);
    m.initialize('c');
    m.doTransform(blocks, reinterpret_cast<unsigned (*)[256]>(freqs));
    for (unsigned k = 0; k < count; ++k) {
      if (!store[k].isTransformed()) throw std::runtime_error("block not flagged transformed");
      nLF_out[k] = (unsigned)store[k].LFpowers().size();
      for (size_t i = 0; i < store[k].LFpowers().size(); ++i) LF_out[k * 256 + i] = store[k].LFpowers()[i];
    }
    return 0;
  } catch (const std::exception& e) {
    if (err && errlen) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return -1;
  }
}

/* sizeof(bwtc_cuda_stats) this library was compiled against (must equal bwtc_cuda_stats_sizeof()). */
unsigned bwtc_host_stats_sizeof(void) { return (unsigned)sizeof(bwtc_cuda_stats); }

int bwtc_host_is_valid_choice(char c) { return bwtc_b200::BWTManager::isValidChoice(c) ? 1 : 0; }

}  // extern "C"
