/* bwtc_b200/host/PipelinedCompressor.cpp — see PipelinedCompressor.hpp.  Compiled against the reference's headers and
 * linked with its objects: every byte of the container is still written by the reference's own PrecompressorBlock,
 * BWTBlock and entropy-coder code; this file only changes WHEN and ON WHICH THREAD that code runs. */
#include "PipelinedCompressor.hpp"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "CudaBWTransform.hpp"
#include "MemStreams.hpp"
#include "PrecompressorBlock.hpp"
#include "RunStatistics.hpp"

namespace bwtc {

namespace {

typedef std::chrono::steady_clock Clock;
double since(const Clock::time_point& t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

struct PbJob;
struct SliceJob {
  PbJob* owner;
  BWTBlock* block;
  MemOutStream out;
  bool encoded;
  SliceJob* chain;  /* CPU engines: the next slice of the same precompressor block, encoded by the same thread right after */
  SliceJob() : owner(0), block(0), encoded(false), chain(0) {}
};
struct PbJob {
  size_t index;
  PrecompressorBlock* pb;
  MemOutStream header;
  std::vector<SliceJob*> slices;
  PbJob() : index(0), pb(0) {}
  ~PbJob() {
    for (size_t i = 0; i < slices.size(); ++i) delete slices[i];
    delete pb;
  }
};

struct Shared {
  std::mutex mu;
  std::condition_variable cvEncode, cvWrite, cvRoom;
  std::deque<SliceJob*> encodeQueue;  /* file order */
  std::deque<PbJob*> writeQueue;      /* file order */
  size_t inFlight;                    /* precompressor blocks read but not yet written */
  bool readerDone, failed;
  std::string error;
  double encoderBusy, bwtWait;
  Shared() : inFlight(0), readerDone(false), failed(false), encoderBusy(0), bwtWait(0) {}
  void fail(const std::string& what) {
    std::lock_guard<std::mutex> g(mu);
    if (!failed) { failed = true; error = what; }
    cvEncode.notify_all();
    cvWrite.notify_all();
    cvRoom.notify_all();
  }
};

void encoderThread(Shared* S, char coderChar, char bwtChoice, uint32 startingPoints) {
  EntropyEncoder* coder = 0;
  try {
    coder = giveEntropyEncoder(coderChar);  /* EntropyCoders.cpp:36-50 */
    BWTManager bwtm;
    bwtm.initialize(bwtChoice);
    bwtm.setStartingPoints(startingPoints);
    double busy = 0;
    for (;;) {
      SliceJob* job = 0;
      {
        std::unique_lock<std::mutex> lk(S->mu);
        S->cvEncode.wait(lk, [&] { return S->failed || !S->encodeQueue.empty() || S->readerDone; });
        if (S->failed) break;
        if (S->encodeQueue.empty()) break;  /* reader done, nothing left */
        job = S->encodeQueue.front();
        S->encodeQueue.pop_front();
      }
      while (job) {
        const Clock::time_point t0 = Clock::now();
        /* the reference's own per-block entry point (HuffmanCoders.cpp:51-61 / WaveletCoders.cpp:80): BWT through the
         * manager — for choice 'c' a wait for the prefetched block — then header, data, back-patched length */
        coder->transformAndEncode(*job->block, bwtm, &job->out);
        runstats::release(job->block->begin());  /* run statistics gathered on the GPU for this block, if any */
        busy += since(t0);
        SliceJob* next = job->chain;
        {
          std::lock_guard<std::mutex> g(S->mu);
          job->encoded = true;
        }
        S->cvWrite.notify_all();
        job = next;
      }
    }
    std::lock_guard<std::mutex> g(S->mu);
    S->encoderBusy += busy;
  } catch (const std::exception& e) {
    S->fail(std::string("encoder thread: ") + e.what());
  }
  delete coder;
}

void writerThread(Shared* S, OutStream* out, bool parts, size_t* written, double* busyOut) {
  double busy = 0;
  try {
    for (;;) {
      PbJob* pj = 0;
      {
        std::unique_lock<std::mutex> lk(S->mu);
        S->cvWrite.wait(lk, [&] {
          if (S->failed) return true;
          if (S->writeQueue.empty()) return S->readerDone;
          PbJob* f = S->writeQueue.front();
          for (size_t i = 0; i < f->slices.size(); ++i) if (!f->slices[i]->encoded) return false;
          return true;
        });
        if (S->failed) break;
        if (S->writeQueue.empty()) break;  /* reader done and everything written */
        pj = S->writeQueue.front();
        S->writeQueue.pop_front();
      }
      const Clock::time_point t0 = Clock::now();
      size_t bytes = pj->header.data().size();
      for (size_t i = 0; i < pj->slices.size(); ++i) bytes += pj->slices[i]->out.data().size();
      if (parts) {  /* part file record: index, length (little endian uint64 each) */
        uint64 hdr[2] = {(uint64)pj->index, (uint64)bytes};
        out->writeBlock(reinterpret_cast<const byte*>(hdr), reinterpret_cast<const byte*>(hdr) + 16);
      }
      const std::vector<byte>& h = pj->header.data();
      if (!h.empty()) out->writeBlock(&h[0], &h[0] + h.size());
      for (size_t i = 0; i < pj->slices.size(); ++i) {
        const std::vector<byte>& d = pj->slices[i]->out.data();
        if (!d.empty()) out->writeBlock(&d[0], &d[0] + d.size());
      }
      *written += bytes;
      delete pj;
      busy += since(t0);
      {
        std::lock_guard<std::mutex> g(S->mu);
        --S->inFlight;
      }
      S->cvRoom.notify_all();
    }
  } catch (const std::exception& e) {
    S->fail(std::string("writer thread: ") + e.what());
  }
  *busyOut = busy;
}

}  // namespace

PipelinedCompressor::PipelinedCompressor(const std::string& in, const std::string& out, const std::string& preprocessing,
                                         size_t memLimit, char entropyCoder)
    : m_in(new BulkFileInStream(in)), m_out(new RawOutStream(out)), m_precompressor(preprocessing),
      m_options(memLimit, entropyCoder), m_bwtChoice('c'), m_startingPoints(1), m_depth(4), m_lookahead(0), m_rank(0),
      m_world(1) { memset(&m_timings, 0, sizeof m_timings); }

PipelinedCompressor::PipelinedCompressor(InStream* in, OutStream* out, const std::string& preprocessing, size_t memLimit,
                                         char entropyCoder)
    : m_in(in), m_out(out), m_precompressor(preprocessing), m_options(memLimit, entropyCoder), m_bwtChoice('c'),
      m_startingPoints(1), m_depth(4), m_lookahead(0), m_rank(0), m_world(1) { memset(&m_timings, 0, sizeof m_timings); }

PipelinedCompressor::~PipelinedCompressor() {
  delete m_in;
  delete m_out;
}

size_t PipelinedCompressor::writeGlobalHeader() {  /* Compressor.cpp:55-58 */
  m_out->writeByte(static_cast<byte>(m_options.entropyCoder));
  return 1;
}

void PipelinedCompressor::initializeBwtAlgorithm(char choice, uint32 startingPoints) {  /* Compressor.cpp:60-63 */
  m_bwtChoice = choice;
  if (startingPoints < 1) startingPoints = 1; else if (startingPoints > 256) startingPoints = 256;  /* BWTManager.cpp:60-64 */
  m_startingPoints = startingPoints;
}

void PipelinedCompressor::setDevices(const std::vector<int>& devices, int depthPerDevice) {
  m_devices = devices;
  m_depth = depthPerDevice < 1 ? 1 : depthPerDevice;
}
void PipelinedCompressor::setLookahead(size_t blocks) { m_lookahead = blocks; }
void PipelinedCompressor::setShard(size_t rank, size_t world) {
  m_world = world < 1 ? 1 : world;
  m_rank = rank % m_world;
}

size_t PipelinedCompressor::compress(size_t threads) {
  if (threads < 1) threads = 1;
  if (m_options.entropyCoder != 'H') threads = 1;  /* wavelet coders carry model state across blocks: encode in order */
  const bool parts = m_world > 1;
  const Clock::time_point tStart = Clock::now();
  memset(&m_timings, 0, sizeof m_timings);
  m_timings.encoderThreads = threads;

  size_t compressedSize = parts ? 0 : writeGlobalHeader();

  /* block sizes exactly as Compressor.cpp:77-81 */
  size_t pbBlockSize = static_cast<size_t>(m_options.memLimit*0.74);
  size_t bwtBlockSize = std::min(static_cast<size_t>(m_options.memLimit*0.185), static_cast<size_t>(0x7fffffff - 1));
  if(m_precompressor.options().size() == 0) pbBlockSize = bwtBlockSize;

  const bool gpu = (m_bwtChoice == 'c');
  /* GPU run statistics for the Huffman coder (BWTC_RUN_STATS=0 turns them off, e.g. for A/B timing) */
  const char* rsEnv = getenv("BWTC_RUN_STATS");
  const bool runStats = gpu && m_options.entropyCoder == 'H' && !(rsEnv && atoi(rsEnv) == 0);
  const size_t gpuSlots = gpu ? std::max<size_t>(1, m_devices.size()) * (size_t)m_depth : 0;
  size_t lookahead = m_lookahead ? m_lookahead : threads + gpuSlots + 2;
  const size_t byBytes = std::max<size_t>(2, (size_t)(8ull << 30) / std::max<size_t>(1, pbBlockSize));
  if (!m_lookahead && lookahead > byBytes) lookahead = byBytes;
  if (lookahead < 2) lookahead = 2;
  if (gpu) CudaBWTransform::configureLookahead(m_devices, m_depth, (uint32)std::min<size_t>(bwtBlockSize, 0x7fffffff - 1));

  Shared S;
  size_t written = 0;
  double writerBusy = 0;
  std::vector<std::thread> encoders;
  for (size_t t = 0; t < threads; ++t)
    encoders.push_back(std::thread(encoderThread, &S, m_options.entropyCoder, m_bwtChoice, m_startingPoints));
  std::thread writer(writerThread, &S, m_out, parts, &written, &writerBusy);

  double readerBusy = 0;
  std::vector<byte> skipBuf;
  try {
    for (size_t index = 0;; ++index) {
      {  /* room in the look-ahead window? */
        std::unique_lock<std::mutex> lk(S.mu);
        S.cvRoom.wait(lk, [&] { return S.failed || S.inFlight < lookahead; });
        if (S.failed) break;
      }
      const Clock::time_point t0 = Clock::now();
      if (parts && index % m_world != m_rank) {  /* another rank's block: consume its raw bytes, keep nothing */
        if (skipBuf.size() < std::min<size_t>(pbBlockSize, 64u << 20)) skipBuf.resize(std::min<size_t>(pbBlockSize, 64u << 20));
        size_t left = pbBlockSize, got = 1;
        bool any = false;
        while (left > 0 && got > 0) {
          got = m_in->readBlock(&skipBuf[0], std::min(left, skipBuf.size()));
          left -= got;
          any = any || got > 0;
        }
        readerBusy += since(t0);
        if (!any) break;  /* end of input */
        continue;
      }
      PrecompressorBlock *pb = m_precompressor.readBlock(pbBlockSize, m_in);  /* Compressor.cpp:86 */
      if(pb->originalSize() == 0) {
        delete pb;
        break;
      }
      if(pbBlockSize != bwtBlockSize) {  /* Compressor.cpp:95-98 */
        size_t s = (m_options.memLimit - pb->size())/4.5;
        bwtBlockSize = std::min(s,static_cast<size_t>(0x7fffffff - 1));
      }
      pb->sliceIntoBlocks(bwtBlockSize);
      PbJob* pj = new PbJob();
      pj->index = index;
      pj->pb = pb;
      pb->writeBlockHeader(&pj->header);  /* Compressor.cpp:104: originalSize, #slices, grammar */
      m_timings.inputBytes += pb->originalSize();
      ++m_timings.precompressorBlocks;
      m_timings.bwtBlocks += pb->slices();
      for(size_t i = 0; i < pb->slices(); ++i) {
        SliceJob* sj = new SliceJob();
        sj->owner = pj;
        sj->block = &pb->getSlice((int)i);
        pj->slices.push_back(sj);
        if (gpu) CudaBWTransform::prefetch(*sj->block, m_startingPoints, runStats);  /* in flight on the GPU from now on */
      }
      readerBusy += since(t0);
      {
        std::lock_guard<std::mutex> g(S.mu);
        ++S.inFlight;
        S.writeQueue.push_back(pj);
        /* CPU engines write one byte past the block while they run (BWTransform.cpp:54-55,63) — the first byte of the
         * next slice — so with them the slices of one precompressor block must not be in flight together: hand them to
         * the encoders one precompressor block at a time (only matters with preprocessing; without it there is one
         * slice per precompressor block, Compressor.cpp:81) */
        if (gpu) {
          for (size_t i = 0; i < pj->slices.size(); ++i) S.encodeQueue.push_back(pj->slices[i]);
        } else {
          for (size_t i = 0; i + 1 < pj->slices.size(); ++i) pj->slices[i]->chain = pj->slices[i + 1];
          if (!pj->slices.empty()) S.encodeQueue.push_back(pj->slices[0]);
        }
      }
      S.cvEncode.notify_all();
      S.cvWrite.notify_all();
    }
  } catch (const std::exception& e) {
    S.fail(std::string("reader: ") + e.what());
  }
  {
    std::lock_guard<std::mutex> g(S.mu);
    S.readerDone = true;
  }
  S.cvEncode.notify_all();
  S.cvWrite.notify_all();
  for (size_t t = 0; t < encoders.size(); ++t) encoders[t].join();
  S.cvWrite.notify_all();
  writer.join();
  /* the per-GPU pipelines stay configured for the next compress() of this process (CudaBWTransform::shutdownLookahead
   * releases them); after a failure they are torn down, which also waits for blocks still in flight */
  if (gpu && S.failed) CudaBWTransform::shutdownLookahead();
  if (S.failed) {
    while (!S.writeQueue.empty()) { delete S.writeQueue.front(); S.writeQueue.pop_front(); }
    throw std::runtime_error("bwtc::PipelinedCompressor: " + S.error);
  }
  compressedSize += written;
  if (!parts) compressedSize += PrecompressorBlock::writeEmptyHeader(m_out);  /* Compressor.cpp:115 */
  m_out->flush();
  m_timings.total = since(tStart);
  m_timings.reader_busy = readerBusy;
  m_timings.encoder_busy_sum = S.encoderBusy;
  m_timings.writer_busy = writerBusy;
  return compressedSize;
}

size_t PipelinedCompressor::mergeParts(const std::vector<std::string>& partFiles, const std::string& outFile, char entropyCoder) {
  struct Rec { uint64 index; size_t file; long offset; uint64 length; };
  std::vector<Rec> recs;
  std::vector<FILE*> files;
  for (size_t f = 0; f < partFiles.size(); ++f) {
    FILE* fp = fopen(partFiles[f].c_str(), "rb");
    if (!fp) throw std::runtime_error("mergeParts: cannot open " + partFiles[f]);
    files.push_back(fp);
    for (;;) {
      uint64 hdr[2];
      if (fread(hdr, 1, 16, fp) != 16) break;
      Rec r = {hdr[0], f, ftell(fp), hdr[1]};
      recs.push_back(r);
      fseek(fp, (long)hdr[1], SEEK_CUR);
    }
  }
  std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.index < b.index; });
  RawOutStream out(outFile);
  out.writeByte(static_cast<byte>(entropyCoder));  /* Compressor.cpp:55-58 */
  size_t total = 1;
  std::vector<byte> buf;
  for (size_t i = 0; i < recs.size(); ++i) {
    if (recs[i].index != i) throw std::runtime_error("mergeParts: precompressor block indices are not 0..n-1");
    buf.resize(recs[i].length);
    fseek(files[recs[i].file], recs[i].offset, SEEK_SET);
    if (recs[i].length && fread(&buf[0], 1, recs[i].length, files[recs[i].file]) != recs[i].length)
      throw std::runtime_error("mergeParts: short read");
    if (recs[i].length) out.writeBlock(&buf[0], &buf[0] + buf.size());
    total += recs[i].length;
  }
  total += PrecompressorBlock::writeEmptyHeader(&out);  /* Compressor.cpp:115 */
  for (size_t f = 0; f < files.size(); ++f) fclose(files[f]);
  return total;
}

} // namespace bwtc
