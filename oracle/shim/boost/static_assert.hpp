/* Test-infrastructure shim for BOOST_STATIC_ASSERT (probmodels/BitPredictors.hpp:33,
 * probmodels/DMC.hpp:35 in the reference).  Not product code. */
#ifndef ORACLE_SHIM_BOOST_STATIC_ASSERT_HPP
#define ORACLE_SHIM_BOOST_STATIC_ASSERT_HPP
#define ORACLE_BSA_CAT2(a, b) a##b
#define ORACLE_BSA_CAT(a, b) ORACLE_BSA_CAT2(a, b)
#define BOOST_STATIC_ASSERT(x) \
  typedef char ORACLE_BSA_CAT(oracle_bsa_, __LINE__)[(x) ? 1 : -1] __attribute__((unused))
#endif
