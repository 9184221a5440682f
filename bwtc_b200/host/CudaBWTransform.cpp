/* bwtc_b200/host/CudaBWTransform.cpp — see CudaBWTransform.hpp.  Compiled against the reference's headers. */
#include "CudaBWTransform.hpp"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

#include "bwtc_cuda.h"
#include "RunStatistics.hpp"

namespace bwtc {

namespace {

int default_device() {
  const char* e = getenv("BWTC_CUDA_DEVICE");
  return e ? atoi(e) : 0;
}

/* Look-ahead state: one bwtc_cuda_pipeline per GPU and the table of blocks in flight, keyed by block address. */
struct Pending {
  bwtc_cuda_pipeline* pipe;
  uint64_t ticket;
  uint32 size;
  uint32 nLF;
  uint32 LF[256];
  uint32 freqs[256];
  bwtc_cuda_runs runs;            /* run statistics asked for with the block (capacity 0: not asked) */
  runstats::Record* runRecord;    /* ... land directly in the record that is published when the block is claimed */
};
struct Lookahead {
  std::mutex mu;
  std::vector<bwtc_cuda_pipeline*> pipes;
  std::vector<int> devices;
  uint32 maxBlock;
  int depth;
  size_t next;
  std::map<const byte*, Pending*> table;
  Lookahead() : maxBlock(0), depth(0), next(0) {}
};
Lookahead& la() {
  static Lookahead L;
  return L;
}

}  // namespace

CudaBWTransform::CudaBWTransform(int device) : m_device(device < 0 ? default_device() : device), m_ctx(0), m_cap(0) {}

CudaBWTransform::~CudaBWTransform() {
  if (m_ctx) bwtc_cuda_ctx_destroy(m_ctx);
}

void CudaBWTransform::fail(const char* what, long long rc) const {
  char buf[768];
  snprintf(buf, sizeof buf, "bwtc::CudaBWTransform::%s failed (%lld): %s", what, rc,
           m_ctx ? bwtc_cuda_last_error(m_ctx) : bwtc_cuda_global_error());
  throw std::runtime_error(buf);
}

void CudaBWTransform::ensure(uint32 block_bytes) const {
  if (m_ctx && block_bytes <= m_cap) return;
  uint64 want = block_bytes < (1u << 20) ? (1u << 20) : block_bytes;
  if (m_ctx) { /* grow geometrically; the old scratch is released first so the peak stays one context */
    if ((uint64)m_cap * 2 > want) want = (uint64)m_cap * 2;
    bwtc_cuda_ctx_destroy(m_ctx);
    m_ctx = 0;
  }
  if (want > BWTC_CUDA_MAX_BLOCK) want = BWTC_CUDA_MAX_BLOCK;
  if (block_bytes > want) fail("ensure (block above the engine limit)", BWTC_CUDA_ETOOBIG);
  int rc = bwtc_cuda_ctx_create(&m_ctx, m_device, (uint32)want);
  if (rc != 0) { m_ctx = 0; fail("ensure/ctx_create", rc); }
  m_cap = (uint32)want;
}

void CudaBWTransform::doTransform(byte *begin, uint32 length, std::vector<uint32>& LF) const {
  ensure(length);
  long long rc = bwtc_cuda_divbwt(m_ctx, begin, begin, length, &LF[0], (uint32)LF.size());
  if (rc < 0) fail("doTransform(raw)", rc);
}

void CudaBWTransform::doTransform(byte *begin, uint32 length, std::vector<uint32>& LF, uint32 freqs[256]) const {
  ensure(length);
  long long rc = bwtc_cuda_divbwtf(m_ctx, begin, begin, length, &LF[0], (uint32)LF.size(), freqs);
  if (rc < 0) fail("doTransform(raw,freqs)", rc);
}

void CudaBWTransform::doTransformFused(BWTBlock& block, uint32 *freqs) const {
  /* a block that was prefetched: wait for it and hand over what the pipeline produced */
  Pending* pd = 0;
  {
    Lookahead& L = la();
    std::lock_guard<std::mutex> g(L.mu);
    std::map<const byte*, Pending*>::iterator it = L.table.find(block.begin());
    if (it != L.table.end()) { pd = it->second; L.table.erase(it); }
  }
  if (pd) {
    const int rc = bwtc_cuda_pipeline_wait(pd->pipe, pd->ticket);
    if (rc < 0) {
      std::string msg = std::string("bwtc::CudaBWTransform: prefetched block failed: ") + bwtc_cuda_pipeline_error(pd->pipe);
      delete pd->runRecord;
      delete pd;
      throw std::runtime_error(msg);
    }
    if (pd->size != block.size() || pd->nLF != block.LFpowers().size()) {
      delete pd->runRecord;
      delete pd;
      throw std::runtime_error("bwtc::CudaBWTransform: block changed between prefetch and doTransform");
    }
    for (uint32 j = 0; j < pd->nLF; ++j) block.LFpowers()[j] = pd->LF[j];
    if (freqs) for (int c = 0; c < 256; ++c) freqs[c] += pd->freqs[c];
    if (pd->runRecord) {
      if (pd->runs.count != BWTC_CUDA_RUNS_OVERFLOW) {  /* few runs: the coder will not have to scan the block */
        pd->runRecord->symbol.resize(pd->runs.count);
        pd->runRecord->start.resize(pd->runs.count);
        runstats::publish(pd->runRecord);
      } else {
        delete pd->runRecord;
      }
    }
    delete pd;
    block.setTransformed(true);
    return;
  }
  ensure((uint32)block.size());
  long long rc = bwtc_cuda_bwt_block(m_ctx, block.begin(), (uint32)block.size(), &block.LFpowers()[0],
                                     (uint32)block.LFpowers().size(), freqs);
  if (rc < 0) fail("doTransformFused(block)", rc);
  block.setTransformed(true);
}

void CudaBWTransform::doTransformFused(std::vector<BWTBlock*>& blocks, uint32 startingPoints, uint32 (*freqs)[256]) const {
  if (blocks.empty()) return;
  /* room for a batch: up to BWTC_CUDA_MAX_BATCH equal-sized small blocks (~32 MiB of text) are one device-side sort */
  uint64 biggest = 0, run = 0;
  for (size_t i = 0; i < blocks.size(); ++i) if (blocks[i]->size() > biggest) biggest = blocks[i]->size();
  for (size_t i = 0; i < blocks.size() && i < BWTC_CUDA_MAX_BATCH; ++i) run += (uint64)blocks[i]->size() + 1;
  uint64 room = run < (32ull << 20) + BWTC_CUDA_MAX_BATCH ? run : (32ull << 20) + BWTC_CUDA_MAX_BATCH;
  if (biggest > room) room = biggest;
  ensure((uint32)(room > BWTC_CUDA_MAX_BLOCK ? BWTC_CUDA_MAX_BLOCK : room));
  std::vector<void*> ptrs(blocks.size());
  std::vector<uint32> sizes(blocks.size()), nLF(blocks.size());
  std::vector<uint32> LF(blocks.size() * 256);
  for (size_t i = 0; i < blocks.size(); ++i) {
    blocks[i]->prepareLFpowers(startingPoints);
    ptrs[i] = blocks[i]->begin();
    sizes[i] = (uint32)blocks[i]->size();
  }
  const int rc = bwtc_cuda_bwt_blocks(m_ctx, &ptrs[0], &sizes[0], (uint32)blocks.size(), startingPoints, 0, &LF[0], &nLF[0],
                                      freqs ? &freqs[0][0] : 0);
  if (rc < 0) fail("doTransformFused(blocks)", rc);
  for (size_t i = 0; i < blocks.size(); ++i) {
    std::vector<uint32>& dst = blocks[i]->LFpowers();
    if (dst.size() != nLF[i]) fail("doTransformFused(blocks): LFpowers sizing differs", BWTC_CUDA_EARG);
    for (uint32 j = 0; j < nLF[i]; ++j) dst[j] = LF[i * 256 + j];
    blocks[i]->setTransformed(true);
  }
}

uint64 CudaBWTransform::maxSizeInBytes(uint64 block_size) const {
  return bwtc_cuda_scratch_bytes((uint32)(block_size > BWTC_CUDA_MAX_BLOCK ? BWTC_CUDA_MAX_BLOCK : block_size));
}
uint64 CudaBWTransform::maxBlockSize(uint64 memory_budget) const {
  if (memory_budget <= BWTC_CUDA_SCRATCH_FIXED_BYTES + 64) return 0;
  uint64 b = (memory_budget - BWTC_CUDA_SCRATCH_FIXED_BYTES) / BWTC_CUDA_SCRATCH_BYTES_PER_SUFFIX - 1;
  return b > BWTC_CUDA_MAX_BLOCK ? BWTC_CUDA_MAX_BLOCK : b;
}
uint64 CudaBWTransform::suggestedBlockSize(uint64 memory_budget) const {
  uint64 b = maxBlockSize(memory_budget);
  return b > (32u << 20) ? (32u << 20) : b;
}

/* ---- inverse ------------------------------------------------------------------------------------------------- */
CudaInverseBWTransform::CudaInverseBWTransform(int device) : m_device(device < 0 ? default_device() : device), m_ctx(0), m_cap(0) {}

CudaInverseBWTransform::~CudaInverseBWTransform() {
  if (m_ctx) bwtc_cuda_ctx_destroy(m_ctx);
}

uint64 CudaInverseBWTransform::maxBlockSize(uint64 memory_budget) const {
  if (memory_budget <= BWTC_CUDA_SCRATCH_FIXED_BYTES + 64) return 0;
  uint64 b = (memory_budget - BWTC_CUDA_SCRATCH_FIXED_BYTES) / BWTC_CUDA_SCRATCH_BYTES_PER_SUFFIX - 1;
  return b > BWTC_CUDA_MAX_BLOCK ? BWTC_CUDA_MAX_BLOCK : b;
}

void CudaInverseBWTransform::doTransform(byte *bwt, uint32 n, const std::vector<uint32>& LFpow) {
  if (n < 2 || LFpow.empty()) return;  /* one row = the end-of-block symbol alone: an empty block */
  if (!m_ctx || n - 1 > m_cap) {
    uint64 want = n - 1 < (1u << 20) ? (1u << 20) : n - 1;
    if (m_ctx) {
      if ((uint64)m_cap * 2 > want) want = (uint64)m_cap * 2;
      bwtc_cuda_ctx_destroy(m_ctx);
      m_ctx = 0;
    }
    if (want > BWTC_CUDA_MAX_BLOCK) want = BWTC_CUDA_MAX_BLOCK;
    const int rc = bwtc_cuda_ctx_create(&m_ctx, m_device, (uint32)want);
    if (rc != 0) {
      m_ctx = 0;
      throw std::runtime_error(std::string("bwtc::CudaInverseBWTransform: cannot create CUDA context: ") + bwtc_cuda_global_error());
    }
    m_cap = (uint32)want;
  }
  const long long rc = bwtc_cuda_inverse_raw(m_ctx, bwt, n, &LFpow[0], (uint32)LFpow.size());
  if (rc < 0) throw std::runtime_error(std::string("bwtc::CudaInverseBWTransform::doTransform failed: ") + bwtc_cuda_last_error(m_ctx));
}

/* ---- look-ahead ---------------------------------------------------------------------------------------------- */
void CudaBWTransform::shutdownLookahead() {
  Lookahead& L = la();
  std::lock_guard<std::mutex> g(L.mu);
  for (std::map<const byte*, Pending*>::iterator it = L.table.begin(); it != L.table.end(); ++it) {
    bwtc_cuda_pipeline_wait(it->second->pipe, it->second->ticket);  /* the engine still writes into those blocks */
    delete it->second->runRecord;
    delete it->second;
  }
  L.table.clear();
  for (size_t i = 0; i < L.pipes.size(); ++i) bwtc_cuda_pipeline_destroy(L.pipes[i]);
  L.pipes.clear();
  L.devices.clear();
  L.maxBlock = 0;
  L.next = 0;
}

void CudaBWTransform::configureLookahead(const std::vector<int>& devices, int depth, uint32 maxBlockBytes) {
  std::vector<int> devs = devices;
  if (devs.empty()) devs.push_back(default_device());
  {
    /* an idle configuration that already fits is kept: creating the per-GPU pipelines (device scratch for `depth`
     * blocks in flight, pinned staging) costs a few hundred milliseconds, a compressor object does not own them */
    Lookahead& L = la();
    std::lock_guard<std::mutex> g(L.mu);
    if (!L.pipes.empty() && L.devices == devs && L.depth == depth && L.maxBlock >= maxBlockBytes && L.table.empty()) return;
  }
  shutdownLookahead();
  Lookahead& L = la();
  std::lock_guard<std::mutex> g(L.mu);
  L.depth = depth;
  for (size_t i = 0; i < devs.size(); ++i) {
    bwtc_cuda_pipeline* p = 0;
    const int rc = bwtc_cuda_pipeline_create(&p, devs[i], depth < 1 ? 1 : depth, maxBlockBytes);
    if (rc != 0) {
      std::string msg = std::string("bwtc::CudaBWTransform::configureLookahead: ") + bwtc_cuda_global_error();
      for (size_t k = 0; k < L.pipes.size(); ++k) bwtc_cuda_pipeline_destroy(L.pipes[k]);
      L.pipes.clear();
      throw std::runtime_error(msg);
    }
    L.pipes.push_back(p);
  }
  L.devices = devs;
  L.maxBlock = maxBlockBytes;
  L.next = 0;
}

void CudaBWTransform::prefetch(BWTBlock& block, uint32 startingPoints, bool wantRunStatistics) {
  Lookahead& L = la();
  std::lock_guard<std::mutex> g(L.mu);
  if (L.pipes.empty() || block.size() == 0 || block.size() > L.maxBlock) return;  /* doTransformFused takes the sync path */
  if (L.table.count(block.begin())) return;
  block.prepareLFpowers(startingPoints);
  Pending* pd = new Pending();
  pd->pipe = L.pipes[L.next++ % L.pipes.size()];
  pd->size = (uint32)block.size();
  pd->nLF = 0;
  for (int c = 0; c < 256; ++c) pd->freqs[c] = 0;
  pd->runRecord = 0;
  pd->runs.capacity = 0;
  int rc;
  /* run statistics: only for blocks that are transformed on their own (small ones are batched on the device), and only
   * worth shipping when the block turns out to have few runs — capacity size/8, the engine reports overflow otherwise */
  if (wantRunStatistics && pd->size > (8u << 20)) {
    pd->runRecord = new runstats::Record();
    pd->runRecord->begin = block.begin();
    pd->runRecord->size = pd->size;
    pd->runRecord->symbol.resize(pd->size / 8);
    pd->runRecord->start.resize(pd->size / 8);
    pd->runs.capacity = pd->size / 8;
    pd->runs.count = BWTC_CUDA_RUNS_OVERFLOW;
    pd->runs.symbol = &pd->runRecord->symbol[0];
    pd->runs.start = &pd->runRecord->start[0];
    rc = bwtc_cuda_pipeline_submit_runs(pd->pipe, block.begin(), block.begin(), pd->size, startingPoints, pd->LF, &pd->nLF, pd->freqs,
                                        &pd->runs, &pd->ticket);
  } else {
    rc = bwtc_cuda_pipeline_submit(pd->pipe, block.begin(), block.begin(), pd->size, startingPoints, 0, pd->LF, &pd->nLF,
                                   pd->freqs, 0, &pd->ticket);
  }
  if (rc < 0) {
    std::string msg = std::string("bwtc::CudaBWTransform::prefetch: ") + bwtc_cuda_pipeline_error(pd->pipe);
    delete pd->runRecord;
    delete pd;
    throw std::runtime_error(msg);
  }
  L.table[block.begin()] = pd;
}

} // namespace bwtc
