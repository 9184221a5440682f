/* bwtc_b200/host/CudaBWTransform.hpp — the B200 forward-BWT engine behind bwtc's OWN operator interface.
 *
 * This header is written against the reference's headers (it is meant to be dropped into bwtransforms/ of a bwtc
 * checkout, see INTEGRATION.md and patches/bwtc-cuda-bwtransform.patch): class bwtc::CudaBWTransform derives from
 * bwtc::BWTransform (bwtransforms/BWTransform.hpp:48-70) exactly like Divsufsorter (bwtransforms/Divsufsorter.hpp:49-72)
 * and SAISBWTransform do — "For implementing new algorithm for Burrows-Wheeler Transform one needs to inherit
 * BWTransform.  After that BWTManager has to be modified" (BWTransform.hpp:44-47).  BWTManager gets one new choice
 * character, 'c'.
 *
 *   raw virtuals  doTransform(byte*, uint32, vector<uint32>&[, freqs])  ->  bwtc_cuda_divbwt / bwtc_cuda_divbwtf
 *                 (what the base-class block wrapper, BWTransform.cpp:39-64, and the reference's tests call)
 *   doTransformFused(BWTBlock&, freqs)   the block-level convention of BWTransform.cpp:52-64 with reverse / sentinel /
 *                 hole fill done on the device (bwtc_cuda_bwt_block); never touches the byte after the block, so
 *                 adjacent slices may be in flight concurrently.  The patched BWTManager::doTransform calls this one.
 *   look-ahead    prefetch(block, startingPoints) queues a block on a per-GPU bwtc_cuda_pipeline and returns at once;
 *                 the later doTransformFused on the same block only waits for it.  This is the "BWTManager prefetch /
 *                 batch extension" a compressor needs to keep several blocks in flight although
 *                 EntropyEncoder::transformAndEncode is synchronous (HuffmanCoders.cpp:51-61, WaveletCoders.cpp:80).
 *
 * Errors: any CUDA failure throws std::runtime_error (the reference silently drops divbwtf's return value,
 * Divsufsorter.hpp:57,64).  There is no CPU fallback.
 */
#ifndef BWTC_CUDA_BWTRANSFORM_HPP_
#define BWTC_CUDA_BWTRANSFORM_HPP_

#include <vector>

#include "BWTransform.hpp"
#include "../globaldefs.hpp"
#include "../BWTBlock.hpp"

struct bwtc_cuda_ctx;

namespace bwtc {

class CudaBWTransform : public BWTransform {
 public:
  /* device < 0: $BWTC_CUDA_DEVICE or 0.  The CUDA context (device scratch) is created on first use and regrown when a
   * larger block arrives. */
  explicit CudaBWTransform(int device = -1);
  virtual ~CudaBWTransform();

  virtual void doTransform(byte *begin, uint32 length, std::vector<uint32>& LF) const;
  virtual void doTransform(byte *begin, uint32 length, std::vector<uint32>& LF, uint32 freqs[256]) const;

  /* Block level, fused on the device.  freqs may be NULL.  LFpowers must have been sized (prepareLFpowers). */
  void doTransformFused(BWTBlock& block, uint32 *freqs) const;
  /* All slices of one precompressor block in one call (the loop of Compressor.cpp:106-109): runs of equal-sized small
   * blocks are sorted as ONE device-side problem (bwtc_cuda_bwt_blocks).  freqs: blocks.size() x 256 or NULL. */
  void doTransformFused(std::vector<BWTBlock*>& blocks, uint32 startingPoints, uint32 (*freqs)[256]) const;

  /* Device scratch a block of this size needs / the largest block that fits a budget (the reference engines return 0
   * here, Divsufsorter.hpp:67-70). */
  virtual uint64 maxSizeInBytes(uint64 block_size) const;
  virtual uint64 maxBlockSize(uint64 memory_budget) const;
  virtual uint64 suggestedBlockSize(uint64 memory_budget) const;

  /* ---- look-ahead (process-wide: the blocks are identified by their address) ------------------------------- */
  /* devices: GPUs to spread blocks over (block i -> devices[i mod G]); depth: blocks in flight per GPU;
   * maxBlockBytes: largest block that will be prefetched.  Replaces any earlier configuration once it is idle. */
  static void configureLookahead(const std::vector<int>& devices, int depth, uint32 maxBlockBytes);
  static void shutdownLookahead();
  /* Queues the in-place transform of `block` (sized by startingPoints as BWTManager would).  The block's bytes, the byte
   * after it excluded, belong to the engine until doTransformFused(block, ...).  wantRunStatistics: also gather the runs of the
   * transformed block on the GPU (RunStatistics.hpp; the reference's TODO at HuffmanCoders.cpp:54) — published when the block
   * is claimed, consumed by the Huffman coder through the wrapped utils::calculateRunFrequenciesAndStoreRuns. */
  static void prefetch(BWTBlock& block, uint32 startingPoints, bool wantRunStatistics = false);

 private:
  void ensure(uint32 block_bytes) const;
  void fail(const char* what, long long rc) const;
  int m_device;
  mutable bwtc_cuda_ctx* m_ctx;
  mutable uint32 m_cap;
};

} // namespace bwtc

#include "InverseBWT.hpp"

namespace bwtc {

/* Inverse transform on the GPU behind the reference's InverseBWTransform interface (bwtransforms/InverseBWT.hpp:45-55),
 * next to MtlSaInverseBWTransform (MtlSaInverseBWT.hpp:42-51).  The patched giveInverseTransformer() (InverseBWT.cpp:42-45
 * takes no argument) returns it when the environment variable BWTC_CUDA_INVERSE is set to a non-zero value.
 * The raw virtual is overridden; the non-virtual block wrapper doTransform(BWTBlock&) (InverseBWT.cpp:47-51) works unchanged
 * on top of it.  Only LFpowers[0] is used: the device makes its own, far denser starting points. */
class CudaInverseBWTransform : public InverseBWTransform {
 public:
  explicit CudaInverseBWTransform(int device = -1);
  virtual ~CudaInverseBWTransform();
  virtual uint64 maxBlockSize(uint64 memory_budget) const;
  virtual void doTransform(byte *bwt, uint32 n, const std::vector<uint32>& LFpow);
 private:
  int m_device;
  bwtc_cuda_ctx* m_ctx;
  uint32 m_cap;
};

} // namespace bwtc

#endif
