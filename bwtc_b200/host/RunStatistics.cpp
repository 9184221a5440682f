/* bwtc_b200/host/RunStatistics.cpp — see RunStatistics.hpp. */
#include "RunStatistics.hpp"

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>

namespace bwtc {
namespace runstats {
namespace {
std::mutex g_mu;
std::map<const byte*, Record*> g_records;  /* keyed by block begin */
std::atomic<size_t> g_count(0), g_served(0);
}  // namespace

void publish(Record* rec) {
  std::lock_guard<std::mutex> g(g_mu);
  std::map<const byte*, Record*>::iterator it = g_records.find(rec->begin);
  if (it != g_records.end()) { delete it->second; g_records.erase(it); --g_count; }
  g_records[rec->begin] = rec;
  ++g_count;
}

void release(const byte* begin) {
  std::lock_guard<std::mutex> g(g_mu);
  std::map<const byte*, Record*>::iterator it = g_records.find(begin);
  if (it != g_records.end()) { delete it->second; g_records.erase(it); --g_count; }
}

size_t served() { return g_served.load(); }

/* The runs of src[0..length) if a record covers that range, else false.  A record's runs are maximal over the whole block;
 * the first and the last run of the range are clipped to it (the reference scans every section on its own, Utils.cpp:150-170,
 * so a run that crosses a section boundary is two runs there). */
bool lookup(uint64* runFreqs, byte* runseq, uint32* runlen, const byte* src, size_t length, uint64* nRuns) {
  if (g_count.load() == 0 || length == 0) return false;
  const Record* rec = 0;
  {
    std::lock_guard<std::mutex> g(g_mu);
    std::map<const byte*, Record*>::iterator it = g_records.upper_bound(src);
    if (it == g_records.begin()) return false;
    --it;
    if (src < it->first || src + length > it->first + it->second->size) return false;
    rec = it->second;  /* stays alive until release(), which the encoding thread calls after it is done with the block */
  }
  const uint32 beg = (uint32)(src - rec->begin), end = beg + (uint32)length;
  /* first run that starts after beg, minus one = the run containing beg */
  size_t k = (size_t)(std::upper_bound(rec->start.begin(), rec->start.end(), beg) - rec->start.begin()) - 1;
  uint64 cnt = 0;
  for (; k < rec->start.size() && rec->start[k] < end; ++k) {
    const uint32 s = std::max(rec->start[k], beg);
    const uint32 e = std::min(k + 1 < rec->start.size() ? rec->start[k + 1] : rec->size, end);
    const byte c = rec->symbol[k];
    ++runFreqs[c];
    runseq[cnt] = c;
    runlen[cnt] = e - s;
    ++cnt;
  }
  *nRuns = cnt;
  ++g_served;
  return true;
}

}  // namespace runstats
}  // namespace bwtc

/* ld --wrap=_ZN5utils35calculateRunFrequenciesAndStoreRunsEPmPhPjPKhm: every call of
 * utils::calculateRunFrequenciesAndStoreRuns in the linked reference objects (HuffmanCoders.cpp:143) lands here. */
extern "C" {
bwtc::uint64 __real__ZN5utils35calculateRunFrequenciesAndStoreRunsEPmPhPjPKhm(bwtc::uint64* runFreqs, bwtc::byte* runseq,
                                                                            bwtc::uint32* runlen, const bwtc::byte* src,
                                                                            size_t length);
bwtc::uint64 __wrap__ZN5utils35calculateRunFrequenciesAndStoreRunsEPmPhPjPKhm(bwtc::uint64* runFreqs, bwtc::byte* runseq,
                                                                            bwtc::uint32* runlen, const bwtc::byte* src,
                                                                            size_t length) {
  bwtc::uint64 n = 0;
  if (bwtc::runstats::lookup(runFreqs, runseq, runlen, src, length, &n)) return n;
  return __real__ZN5utils35calculateRunFrequenciesAndStoreRunsEPmPhPjPKhm(runFreqs, runseq, runlen, src, length);
}
}
