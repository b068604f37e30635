/* ref_shim.h -- force-included (-include) in front of the UNMODIFIED reference sources when the
 * shared-RNG reference oracle is built (oracle/Makefile, target _ref/libref_cb.so).
 *
 * TEST INFRASTRUCTURE: see oracle/README.md.  Nothing here is linked into the product.
 *
 * It swaps the reference's global engine type (std::default_random_engine,
 * /root/reference/source/LightSource.h:6) for a 32-bit UniformRandomBitGenerator that returns the
 * next word of the counter-based stream in oracle/rng_contract.h.  The reference sources are not
 * edited: the swap is a macro on the type name, applied after <random> itself has been parsed.
 */
#ifndef RT_ORACLE_REF_SHIM_H
#define RT_ORACLE_REF_SHIM_H
#include <random>
#include <cstdint>
namespace std {
struct cb_engine {
  typedef uint32_t result_type;
  static constexpr result_type min() { return 0u; }
  static constexpr result_type max() { return 0xFFFFFFFFu; }
  result_type operator()();          /* defined in ref_harness.cpp */
};
}  // namespace std
#define default_random_engine cb_engine
#define RT_ORACLE_SHARED_RNG 1
#endif
