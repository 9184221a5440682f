/* bwtc_b200/host/RunStatistics.hpp — run statistics gathered during the BWT.
 *
 * HuffmanEncoder::encodeData splits every section of the transformed block into runs by scanning it byte by byte
 * (utils::calculateRunFrequenciesAndStoreRuns, Utils.cpp:150-170) and the reference notes "TODO: Also gather information
 * about the runs during BWT" (HuffmanCoders.cpp:54).  The GPU engine can emit the maximal runs of the whole block
 * (bwtc_cuda_runs, include/bwtc_cuda.h).  This file is the host side: a registry of per-block run records and a
 * link-time wrapper (ld --wrap) around utils::calculateRunFrequenciesAndStoreRuns that answers from the registry —
 * slicing the block's runs at the section boundaries the coder asks for — and falls back to the reference's scan when
 * no record covers the range.  The coder's sources are not touched and the bytes written are the same either way.
 */
#ifndef BWTC_B200_RUN_STATISTICS_HPP_
#define BWTC_B200_RUN_STATISTICS_HPP_

#include <vector>

#include "globaldefs.hpp"

namespace bwtc {
namespace runstats {

struct Record {
  const byte* begin;           /* the transformed block these runs describe ... */
  uint32 size;                 /* ... and its length */
  std::vector<byte> symbol;    /* symbol[k] */
  std::vector<uint32> start;   /* start[k]; run k ends at start[k+1] (the last one at size) */
};

void publish(Record* rec);            /* takes ownership; replaces an older record of the same block */
void release(const byte* begin);      /* the block has been encoded */
size_t served();                      /* calls answered from the registry so far (tests, timing reports) */

}  // namespace runstats
}  // namespace bwtc

#endif
