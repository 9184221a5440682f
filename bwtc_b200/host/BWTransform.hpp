/* bwtc_b200/host/BWTransform.hpp — C++ host-side mirror of the reference's operator interface for the
 * forward-BWT path, above the C-ABI of include/bwtc_cuda.h.  Same class names, method names, argument
 * meaning and error behaviour as the reference, so bwtc's Compressor / entropy coders can use it unchanged:
 *
 *   bwtc_b200::BWTBlock         <-> bwtc::BWTBlock           (BWTBlock.hpp:39-72, BWTBlock.cpp:104-108)
 *   bwtc_b200::BWTransform      <-> bwtc::BWTransform        (bwtransforms/BWTransform.hpp:48-70)
 *   bwtc_b200::CudaBWTransform  :   the new subclass ("inherit BWTransform, then modify BWTManager",
 *                                   bwtransforms/BWTransform.hpp:44-47) — replaces Divsufsorter
 *                                   (bwtransforms/Divsufsorter.hpp:49-72) and SAISBWTransform
 *   bwtc_b200::BWTManager       <-> bwtc::BWTManager         (bwtransforms/BWTManager.hpp:40-58) + choice 'c'
 *
 * Differences, all deliberate:
 *   - doTransform(BWTBlock&[, freqs]) is VIRTUAL here (non-virtual in the reference, BWTransform.hpp:60-61) so
 *     the CUDA subclass can fuse reverse / sentinel / hole-fill on the device instead of running
 *     std::reverse on the host (BWTransform.cpp:53); the base-class implementation keeps the reference's
 *     exact host-side behaviour on top of the raw virtual.
 *   - any CUDA failure throws std::runtime_error (the reference silently drops divbwtf's return value,
 *     Divsufsorter.hpp:57,64).  There is no CPU fallback.
 */
#ifndef BWTC_B200_HOST_BWTRANSFORM_HPP_
#define BWTC_B200_HOST_BWTRANSFORM_HPP_

#include <stdint.h>

#include <cassert>
#include <cstddef>
#include <vector>

#include "../../include/bwtc_cuda.h"

namespace bwtc_b200 {

typedef unsigned char byte;
typedef uint32_t uint32;
typedef uint64_t uint64;

class BWTBlock {
 public:
  BWTBlock() : m_begin(0), m_length(0), m_isTransformed(false) {}
  BWTBlock(byte* data, uint32 length, bool isTransformed)
      : m_begin(data), m_length(length), m_isTransformed(isTransformed) {}
  void setTransformed(bool transformed) { assert(m_isTransformed != transformed); m_isTransformed = transformed; }
  bool isTransformed() const { return m_isTransformed; }
  size_t size() const { return m_length; }
  byte* begin() { return m_begin; }
  const byte* begin() const { return m_begin; }
  byte* end() { return m_begin + m_length; }
  const byte* end() const { return m_begin + m_length; }
  std::vector<uint32>& LFpowers() { return m_LFpowers; }
  void setBegin(byte* begin) { m_begin = begin; }
  void setSize(uint32 length) { m_length = length; }
  void prepareLFpowers(uint32 startingPoints);  /* BWTBlock.cpp:104-108 */

 private:
  byte* m_begin;
  uint32 m_length;
  std::vector<uint32> m_LFpowers;
  bool m_isTransformed;
};

class BWTransform {
 public:
  BWTransform() {}
  virtual ~BWTransform() {}
  /* raw contract: begin holds `length` bytes, normally reverse(block) + 0x00 (BWTransform.hpp:53-58) */
  virtual void doTransform(byte* begin, uint32 length, std::vector<uint32>& LF) const = 0;
  virtual void doTransform(byte* begin, uint32 length, std::vector<uint32>& LF, uint32 freqs[256]) const = 0;
  /* block contract (BWTransform.cpp:39-64) */
  virtual void doTransform(BWTBlock& block);
  virtual void doTransform(BWTBlock& block, uint32 freqs[256]);
  virtual uint64 maxSizeInBytes(uint64 block_size) const = 0;
  virtual uint64 maxBlockSize(uint64 memory_budget) const = 0;
  virtual uint64 suggestedBlockSize(uint64 memory_budget) const = 0;

 private:
  BWTransform(const BWTransform&);
  const BWTransform& operator=(const BWTransform&);
};

class CudaBWTransform : public BWTransform {
 public:
  explicit CudaBWTransform(int device = 0, uint32 initial_max_block = 1u << 20);
  virtual ~CudaBWTransform();
  virtual void doTransform(byte* begin, uint32 length, std::vector<uint32>& LF) const;
  virtual void doTransform(byte* begin, uint32 length, std::vector<uint32>& LF, uint32 freqs[256]) const;
  virtual void doTransform(BWTBlock& block);                      /* fused on the device */
  virtual void doTransform(BWTBlock& block, uint32 freqs[256]);   /* fused on the device */
  /* All slices of one precompressor block in one call (the loop of Compressor.cpp:100-109): runs of equal-sized
   * small blocks are sorted as ONE device-side problem (bwtc_cuda_bwt_blocks).  freqs: blocks.size() x 256 counters
   * (incremented) or NULL.  LFpowers of every block must have been sized (prepareLFpowers) by the caller. */
  void doTransform(std::vector<BWTBlock*>& blocks, uint32 starts, uint32 (*freqs)[256]);
  virtual uint64 maxSizeInBytes(uint64 block_size) const;
  virtual uint64 maxBlockSize(uint64 memory_budget) const;
  virtual uint64 suggestedBlockSize(uint64 memory_budget) const;
  const bwtc_cuda_stats& lastStats() const { return m_stats; }

 private:
  void ensure(uint32 block_bytes) const;
  void fail(const char* what, long long rc) const;
  int m_device;
  mutable bwtc_cuda_ctx* m_ctx;
  mutable uint32 m_cap;
  mutable bwtc_cuda_stats m_stats;
};

/* 'c' -> CudaBWTransform.  The reference's 'd' / 's' / 'a' CPU engines are not carried: asking for them
 * throws (no fallback).  giveTransformer mirrors bwtransforms/BWTransform.cpp:66-76. */
BWTransform* giveTransformer(char transform);

class BWTManager {
 public:
  BWTManager();
  explicit BWTManager(uint32 startingPoints);
  ~BWTManager();
  void doTransform(BWTBlock& block);
  void doTransform(BWTBlock& block, uint32* freqs);
  /* batched look-ahead extension: every block gets exactly what doTransform(block, freqs)
   * would give it; freqs = blocks.size() x 256 counters or NULL */
  void doTransform(std::vector<BWTBlock*>& blocks, uint32 (*freqs)[256]);
  void initialize(char choice);
  void setStartingPoints(uint32 startingPoints);
  uint32 getStartingPoints() const;
  static bool isValidChoice(char c);

 private:
  std::vector<BWTransform*> m_transformers;
  uint32 m_startingPoints;
};

}  // namespace bwtc_b200

#endif
