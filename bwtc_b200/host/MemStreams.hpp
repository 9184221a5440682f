/* bwtc_b200/host/MemStreams.hpp — in-memory / bulk-read implementations of the reference's stream interfaces
 * (Streams.hpp:40-60) for the pipelined compressor.
 *
 *   MemOutStream      OutStream over a byte vector.  One per BWT block: HuffmanEncoder back-patches the 48-bit block
 *                     length at getPos() (HuffmanCoders.cpp:259-261,275), which is block-relative here, so blocks can be
 *                     encoded on any thread and concatenated in file order afterwards.  Same idea as the reference's own
 *                     test backend (test/TestStreams.hpp:38-125), written for this purpose.
 *   BulkFileInStream  InStream whose readBlock is one fread (RawInStream::readBlock copies byte by byte through
 *                     fetchByte(), Streams.cpp:146-154 — about 0.3 GB/s, far below what the GPU stage consumes).
 *   MemInStream       InStream over caller memory (bench: synthetic input without a file system in the timed region).
 * Only readBlock is needed by the compressor side (PrecompressorBlock.cpp:41-49); the bit-level readers abort.
 */
#ifndef BWTC_B200_MEMSTREAMS_HPP_
#define BWTC_B200_MEMSTREAMS_HPP_

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <stdint.h>
#if defined(__linux__)
#include <sys/mman.h>
#endif

#include "Streams.hpp"

namespace bwtc {

/* PrecompressorBlock hands readBlock a freshly malloc'ed (mmap'ed, untouched) buffer of the whole block size
 * (PrecompressorBlock.cpp:41-45): filling it is one page fault per 4 KiB — measured 0.33 GB/s for the reader thread,
 * which would starve the whole pipeline.  Asking for transparent huge pages on the buffer's interior before the first
 * touch makes that one fault per 2 MiB (1.3 GB/s here).  A hint only: harmless where THP is off. */
inline void adviseHugePages(byte *to, size_t size) {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
  const uintptr_t huge = (uintptr_t)2 << 20;
  if(size < 2 * huge) return;
  const uintptr_t lo = ((uintptr_t)to + huge - 1) & ~(huge - 1), hi = ((uintptr_t)to + size) & ~(huge - 1);
  if(hi > lo) madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
#else
  (void)to; (void)size;
#endif
}

class MemOutStream : public OutStream {
 public:
  MemOutStream() {}
  virtual ~MemOutStream() {}
  virtual void writeByte(byte b) { m_data.push_back(b); }
  virtual void writeBlock(const byte *begin, const byte *end) { m_data.insert(m_data.end(), begin, end); }
  virtual long int getPos() { return (long int)m_data.size(); }
  virtual void write48bits(uint64 to_written, long int position) {
    for(int i = 5; i >= 0; --i) m_data[position++] = (byte)(0xFF & (to_written >> i*8));
  }
  virtual void flush() {}
  std::vector<byte>& data() { return m_data; }
  const std::vector<byte>& data() const { return m_data; }
 private:
  std::vector<byte> m_data;
};

class InStreamReadOnlyBlocks : public InStream {
 public:
  virtual bool readBit() { unsupported(); return false; }
  virtual byte readByte() { unsupported(); return 0; }
  virtual void flushBuffer() {}
  virtual uint64 read48bits() { unsupported(); return 0; }
 private:
  static void unsupported() {
    fprintf(stderr, "bwtc_b200: this InStream only supports readBlock (compressor side)\n");
    abort();
  }
};

class BulkFileInStream : public InStreamReadOnlyBlocks {
 public:
  explicit BulkFileInStream(const std::string& name) : m_file(name == "" ? stdin : fopen(name.c_str(), "rb")) {
    if(!m_file) { perror(name.c_str()); exit(1); }  /* same behaviour as RawInStream, Streams.cpp:126-131 */
  }
  virtual ~BulkFileInStream() { if(m_file != stdin) fclose(m_file); }
  virtual size_t readBlock(byte *to, size_t max_block_size) {
    adviseHugePages(to, max_block_size);
    size_t have = 0;
    while(have < max_block_size) {
      size_t r = fread(to + have, 1, max_block_size - have, m_file);
      if(r == 0) break;
      have += r;
    }
    return have;
  }
  virtual bool compressedDataEnding() { return feof(m_file) != 0; }
 private:
  FILE *m_file;
};

class MemInStream : public InStreamReadOnlyBlocks {
 public:
  MemInStream(const byte *data, size_t size) : m_data(data), m_size(size), m_pos(0) {}
  virtual size_t readBlock(byte *to, size_t max_block_size) {
    size_t r = m_size - m_pos < max_block_size ? m_size - m_pos : max_block_size;
    adviseHugePages(to, r);
    memcpy(to, m_data + m_pos, r);
    m_pos += r;
    return r;
  }
  virtual bool compressedDataEnding() { return m_pos >= m_size; }
 private:
  const byte *m_data;
  size_t m_size, m_pos;
};

/* OutStream that forwards to memory the caller reads afterwards (bench / tests), or only counts. */
class CountingOutStream : public OutStream {
 public:
  CountingOutStream() : m_count(0) {}
  virtual void writeByte(byte) { ++m_count; }
  virtual void writeBlock(const byte *begin, const byte *end) { m_count += end - begin; }
  virtual long int getPos() { return (long int)m_count; }
  virtual void write48bits(uint64, long int) {}
  virtual void flush() {}
  uint64 count() const { return m_count; }
 private:
  uint64 m_count;
};

} // namespace bwtc

#endif
