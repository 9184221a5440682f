/* bwtc_b200/host/BWTManager.cpp — mirror of bwtransforms/BWTManager.cpp:35-80 with the one new choice 'c'. */
#include <stdexcept>

#include "BWTransform.hpp"

namespace bwtc_b200 {

BWTManager::BWTManager() : m_startingPoints(1) {}
BWTManager::BWTManager(uint32 startingPoints) : m_startingPoints(startingPoints) {}

BWTManager::~BWTManager() {
  for (size_t i = 0; i < m_transformers.size(); ++i) delete m_transformers[i];
}

void BWTManager::doTransform(BWTBlock& block) {  /* BWTManager.cpp:46-51 */
  assert(!block.isTransformed());
  block.prepareLFpowers(m_startingPoints);
  m_transformers.at(0)->doTransform(block);
}

void BWTManager::doTransform(BWTBlock& block, uint32* freqs) {  /* BWTManager.cpp:53-58 */
  assert(!block.isTransformed());
  block.prepareLFpowers(m_startingPoints);
  m_transformers.at(0)->doTransform(block, freqs);
}

void BWTManager::doTransform(std::vector<BWTBlock*>& blocks, uint32 (*freqs)[256]) {
  for (size_t i = 0; i < blocks.size(); ++i) {
    assert(!blocks[i]->isTransformed());
    blocks[i]->prepareLFpowers(m_startingPoints);
  }
  CudaBWTransform* t = dynamic_cast<CudaBWTransform*>(m_transformers.at(0));
  if (!t) throw std::logic_error("bwtc_b200::BWTManager: batched doTransform needs the CUDA transformer");
  t->doTransform(blocks, m_startingPoints, freqs);
}

void BWTManager::setStartingPoints(uint32 startingPoints) {  /* BWTManager.cpp:60-64 */
  if (startingPoints < 1) startingPoints = 1;
  else if (startingPoints > 256) startingPoints = 256;
  m_startingPoints = startingPoints;
}

uint32 BWTManager::getStartingPoints() const { return m_startingPoints; }

bool BWTManager::isValidChoice(char c) { return c == 'c'; }

void BWTManager::initialize(char choice) {  /* BWTManager.cpp:74-80 */
  if (!isValidChoice(choice))
    throw std::invalid_argument("bwtc_b200::BWTManager::initialize: only choice 'c' (CUDA) exists");
  m_transformers.push_back(new CudaBWTransform());
}

}  // namespace bwtc_b200
