/* bwtc_b200/host/PipelinedCompressor.hpp — batched look-ahead replacement for bwtc::Compressor (Compressor.hpp:99-121,
 * Compressor.cpp:65-120) that writes the SAME bytes.
 *
 * The reference loop is: read a precompressor block -> slice it -> for every slice transformAndEncode (BWT, then entropy
 * coding, synchronously) -> next block.  One block is in flight, one core is busy, the GPU would idle while Huffman codes
 * and vice versa.  Here the three stages overlap:
 *
 *   reader (caller thread)   Precompressor::readBlock + sliceIntoBlocks, exactly as Compressor.cpp:86-104; every slice is
 *                            prefetched on the GPU pipeline at once (CudaBWTransform::prefetch) — up to `lookahead`
 *                            precompressor blocks ahead of the writer
 *   encoder threads          each owns an EntropyEncoder (giveEntropyEncoder) and a BWTManager; takes the next slice in
 *                            file order and calls the reference's own transformAndEncode on it with a per-slice in-memory
 *                            OutStream — the BWT inside it is just a wait for the prefetched result.  Coder 'H' re-initialises
 *                            all its state per block (HuffmanCoders.cpp:274-275,312), so slices are encoded in parallel;
 *                            the wavelet coders carry predictor state from block to block (probmodels/FSM.hpp:205-218), so
 *                            for them ONE encoder thread runs, in order — the GPU look-ahead still applies
 *   writer thread            precompressor-block header (PrecompressorBlock::writeBlockHeader), then the slices' streams in
 *                            order; finally PrecompressorBlock::writeEmptyHeader
 *
 * The output is byte-identical to Compressor::compress for every coder, BWT choice and preprocessing string (the
 * bytes of a block depend only on (BWT bytes, LFpowers, freqs) and the coder state).
 * With BWT choice 'd' / 's' the reference's CPU engines run inside the encoder threads (that is how the ordering logic
 * is tested without a GPU); choice 'c' is the product path.
 */
#ifndef BWTC_B200_PIPELINED_COMPRESSOR_HPP_
#define BWTC_B200_PIPELINED_COMPRESSOR_HPP_

#include <string>
#include <vector>

#include "bwtransforms/BWTManager.hpp"
#include "preprocessors/Precompressor.hpp"
#include "EntropyCoders.hpp"
#include "Streams.hpp"
#include "Compressor.hpp" /* Options */

namespace bwtc {

struct PipelineTimings {  /* seconds, filled by compress() */
  double total, reader_busy, encoder_busy_sum, writer_busy, bwt_wait_sum;
  size_t precompressorBlocks, bwtBlocks, inputBytes, encoderThreads;
};

class PipelinedCompressor {
 public:
  PipelinedCompressor(const std::string& in, const std::string& out, const std::string& preprocessing, size_t memLimit,
                      char entropyCoder);
  PipelinedCompressor(InStream* in, OutStream* out, const std::string& preprocessing, size_t memLimit, char entropyCoder);
  ~PipelinedCompressor();

  /* same three calls as Compressor (compress.cpp:192-195) */
  void initializeBwtAlgorithm(char choice, uint32 startingPoints);
  size_t writeGlobalHeader();
  /* threads = host threads for entropy coding (>= 1; forced to 1 for coders other than 'H').  Returns compressed size. */
  size_t compress(size_t threads);

  /* GPU side (choice 'c'): devices to spread blocks over, blocks in flight per device */
  void setDevices(const std::vector<int>& devices, int depthPerDevice);
  /* precompressor blocks the reader may run ahead of the writer (memory = lookahead x block size); 0 = automatic */
  void setLookahead(size_t blocks);
  /* only precompressor blocks with index % world == rank are read into memory, transformed and encoded; the output then
   * is a part file: { uint64 index, uint64 length, bytes }* of this rank's blocks (no global header, no terminator) that
   * mergeParts() interleaves with the other ranks' — one process per GPU, no data exchange before the final concat */
  void setShard(size_t rank, size_t world);
  static size_t mergeParts(const std::vector<std::string>& partFiles, const std::string& outFile, char entropyCoder);

  const PipelineTimings& timings() const { return m_timings; }

 private:
  InStream *m_in;
  OutStream *m_out;
  Precompressor m_precompressor;
  Options m_options;
  char m_bwtChoice;
  uint32 m_startingPoints;
  std::vector<int> m_devices;
  int m_depth;
  size_t m_lookahead, m_rank, m_world;
  PipelineTimings m_timings;

  PipelinedCompressor(const PipelinedCompressor&);
  PipelinedCompressor& operator=(const PipelinedCompressor&);
};

} // namespace bwtc

#endif
