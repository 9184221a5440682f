/* bwtc_b200/host/CudaBWTransform.cpp — see BWTransform.hpp. */
#include <algorithm>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "BWTransform.hpp"

namespace bwtc_b200 {

void BWTBlock::prepareLFpowers(uint32 startingPoints) {  /* BWTBlock.cpp:104-108 */
  if (m_length <= 256 || startingPoints == 0) m_LFpowers.resize(1);
  else if (startingPoints <= 256) m_LFpowers.resize(startingPoints);
  else m_LFpowers.resize(256);
}

/* Base-class block wrappers: byte-for-byte the reference's host-side convention (BWTransform.cpp:39-64). */
void BWTransform::doTransform(BWTBlock& block) {
  std::reverse(block.begin(), block.end());
  byte next = *block.end();
  *block.end() = 0;
  doTransform(block.begin(), (uint32)block.size() + 1, block.LFpowers());
  block.setTransformed(true);
  *(block.begin() + block.LFpowers()[0]) = *block.end();
  *block.end() = next;
}

void BWTransform::doTransform(BWTBlock& block, uint32 freqs[256]) {
  std::reverse(block.begin(), block.end());
  byte next = *block.end();
  *block.end() = 0;
  doTransform(block.begin(), (uint32)block.size() + 1, block.LFpowers(), freqs);
  block.setTransformed(true);
  *(block.begin() + block.LFpowers()[0]) = *block.end();
  *block.end() = next;
}

CudaBWTransform::CudaBWTransform(int device, uint32 initial_max_block)
    : m_device(device), m_ctx(0), m_cap(0) {
  ensure(initial_max_block);
}

CudaBWTransform::~CudaBWTransform() { bwtc_cuda_ctx_destroy(m_ctx); }

void CudaBWTransform::fail(const char* what, long long rc) const {
  char buf[768];
  snprintf(buf, sizeof buf, "CudaBWTransform::%s failed (%lld): %s", what, rc,
           m_ctx ? bwtc_cuda_last_error(m_ctx) : bwtc_cuda_global_error());
  throw std::runtime_error(buf);
}

void CudaBWTransform::ensure(uint32 block_bytes) const {
  if (m_ctx && block_bytes <= m_cap) return;
  uint32 want = block_bytes;
  if (m_ctx) {  /* grow geometrically, like a vector */
    uint64 g = (uint64)m_cap * 2;
    if (g > BWTC_CUDA_MAX_BLOCK) g = BWTC_CUDA_MAX_BLOCK;
    if (g > want) want = (uint32)g;
    bwtc_cuda_ctx_destroy(m_ctx);
    m_ctx = 0;
  }
  int rc = bwtc_cuda_ctx_create(&m_ctx, m_device, want);
  if (rc != 0) { m_ctx = 0; fail("ensure/ctx_create", rc); }
  m_cap = want;
}

void CudaBWTransform::doTransform(byte* begin, uint32 length, std::vector<uint32>& LF) const {
  ensure(length);
  long long rc = bwtc_cuda_divbwt(m_ctx, begin, begin, length, &LF[0], (uint32)LF.size());
  if (rc < 0) fail("doTransform(raw)", rc);
  bwtc_cuda_get_stats(m_ctx, &m_stats);
}

void CudaBWTransform::doTransform(byte* begin, uint32 length, std::vector<uint32>& LF, uint32 freqs[256]) const {
  ensure(length);
  long long rc = bwtc_cuda_divbwtf(m_ctx, begin, begin, length, &LF[0], (uint32)LF.size(), freqs);
  if (rc < 0) fail("doTransform(raw,freqs)", rc);
  bwtc_cuda_get_stats(m_ctx, &m_stats);
}

void CudaBWTransform::doTransform(BWTBlock& block) {
  ensure((uint32)block.size());
  long long rc = bwtc_cuda_bwt_block(m_ctx, block.begin(), (uint32)block.size(), &block.LFpowers()[0],
                                     (uint32)block.LFpowers().size(), 0);
  if (rc < 0) fail("doTransform(block)", rc);
  block.setTransformed(true);
  bwtc_cuda_get_stats(m_ctx, &m_stats);
}

void CudaBWTransform::doTransform(BWTBlock& block, uint32 freqs[256]) {
  ensure((uint32)block.size());
  long long rc = bwtc_cuda_bwt_block(m_ctx, block.begin(), (uint32)block.size(), &block.LFpowers()[0],
                                     (uint32)block.LFpowers().size(), freqs);
  if (rc < 0) fail("doTransform(block,freqs)", rc);
  block.setTransformed(true);
  bwtc_cuda_get_stats(m_ctx, &m_stats);
}

void CudaBWTransform::doTransform(std::vector<BWTBlock*>& blocks, uint32 starts, uint32 (*freqs)[256]) {
  if (blocks.empty()) return;
  /* room for a batch: up to 64 equal-sized small blocks (~32 MiB of text) are sorted as one device-side problem */
  uint64 biggest = 0, run = 0;
  for (size_t i = 0; i < blocks.size(); ++i) biggest = std::max<uint64>(biggest, blocks[i]->size());
  for (size_t i = 0; i < blocks.size() && i < BWTC_CUDA_MAX_BATCH; ++i) run += (uint64)blocks[i]->size() + 1;
  const uint64 batch_room = std::min<uint64>(run, (32ull << 20) + BWTC_CUDA_MAX_BATCH);
  ensure((uint32)std::min<uint64>(std::max<uint64>(biggest, batch_room), BWTC_CUDA_MAX_BLOCK));
  std::vector<void*> ptrs(blocks.size());
  std::vector<uint32> sizes(blocks.size()), nLF(blocks.size());
  std::vector<uint32> LF(blocks.size() * 256);
  for (size_t i = 0; i < blocks.size(); ++i) { ptrs[i] = blocks[i]->begin(); sizes[i] = (uint32)blocks[i]->size(); }
  const int rc = bwtc_cuda_bwt_blocks(m_ctx, &ptrs[0], &sizes[0], (uint32)blocks.size(), starts, 0, &LF[0], &nLF[0],
                                      freqs ? &freqs[0][0] : 0);
  if (rc < 0) fail("doTransform(blocks)", rc);
  for (size_t i = 0; i < blocks.size(); ++i) {
    std::vector<uint32>& dst = blocks[i]->LFpowers();
    if (dst.size() != nLF[i]) fail("doTransform(blocks): LFpowers not sized by prepareLFpowers", BWTC_CUDA_EARG);
    for (uint32 j = 0; j < nLF[i]; ++j) dst[j] = LF[i * 256 + j];
    blocks[i]->setTransformed(true);
  }
  bwtc_cuda_get_stats(m_ctx, &m_stats);
}

/* Unlike the reference engines (which return 0, Divsufsorter.hpp:67-70) these are real: ~37 bytes of device
 * scratch per suffix plus look-back status words. */
uint64 CudaBWTransform::maxSizeInBytes(uint64 block_size) const { return 38 * (block_size + 1) + (1u << 20); }
uint64 CudaBWTransform::maxBlockSize(uint64 memory_budget) const {
  if (memory_budget <= (1u << 20) + 64) return 0;
  uint64 b = (memory_budget - (1u << 20)) / 38 - 1;
  return b > BWTC_CUDA_MAX_BLOCK ? BWTC_CUDA_MAX_BLOCK : b;
}
uint64 CudaBWTransform::suggestedBlockSize(uint64 memory_budget) const {
  uint64 b = maxBlockSize(memory_budget);
  return b > (32u << 20) ? (32u << 20) : b;
}

BWTransform* giveTransformer(char transform) {
  if (transform != 'c')
    throw std::invalid_argument("bwtc_b200::giveTransformer: only the CUDA transformer 'c' exists (no CPU fallback)");
  return new CudaBWTransform();
}

}  // namespace bwtc_b200
