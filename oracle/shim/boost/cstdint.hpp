/* Test-infrastructure shim: lets the UNMODIFIED reference sources (which include
 * <boost/cstdint.hpp> only for fixed-width typedefs, globaldefs.hpp:31,44-50)
 * compile in an image without Boost.  Not product code. */
#ifndef ORACLE_SHIM_BOOST_CSTDINT_HPP
#define ORACLE_SHIM_BOOST_CSTDINT_HPP
#include <stdint.h>
namespace boost {
using ::int8_t;  using ::uint8_t;
using ::int16_t; using ::uint16_t;
using ::int32_t; using ::uint32_t;
using ::int64_t; using ::uint64_t;
}
#endif
