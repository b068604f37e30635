/* rng_contract.h -- the counter-based random stream shared by the oracle and the CUDA kernels.
 *
 * TEST INFRASTRUCTURE (oracle/): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything in this directory.  The product re-states this contract
 * independently in ray-tracing-engine_b200/csrc/rng.cuh; DESIGN.md "RNG contract" is the normative text.
 *
 * The reference draws every random number from one global std::default_random_engine
 * (/root/reference/source/LightSource.h:6) consumed in serial order.  To make per-sample parity
 * possible both sides instead consume, for every independent unit of
 * work (one pixel sample, one photon path), the words w(0), w(1), w(2), ... of a stream keyed by
 * (seed, domain, index):
 *
 *   mix(z)      : z ^= z>>30; z *= 0xBF58476D1CE4E5B9; z ^= z>>27; z *= 0x94D049BB133111EB; z ^= z>>31
 *   K           = mix( mix(seed + GOLDEN) ^ ((domain << 56) | index) )
 *   pair(j)     = mix( K + (j+1) * GOLDEN )                 GOLDEN = 0x9E3779B97F4A7C15
 *   w(c)        = c even ? low 32 bits of pair(c/2) : high 32 bits of pair(c/2)
 *
 * and turn words into uniforms exactly the way libstdc++'s generate_canonical does for a 32-bit
 * engine (bits/random.tcc: generate_canonical):
 *   float  canonical : f = float(w) / 2^32 ; if (f >= 1) f = nextafterf(1, 0)
 *   double canonical : g = (double(w0) + double(w1) * 2^32) / 2^64      (w0 drawn first)
 *   uniform_real_distribution(a, b) : canonical * (b - a) + a   (no fused multiply-add)
 */
#ifndef RT_ORACLE_RNG_CONTRACT_H
#define RT_ORACLE_RNG_CONTRACT_H
#include <stdint.h>

#define RTO_GOLDEN 0x9E3779B97F4A7C15ull
#define RTO_DOMAIN_PIXEL 1ull   /* index = sample * (W*H) + y*W + x                      */
#define RTO_DOMAIN_PHOTON 2ull  /* index = light * photons_per_light + path              */

static inline uint64_t rto_mix(uint64_t z) {
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31; return z;
}
static inline uint64_t rto_stream_key(uint64_t seed, uint64_t domain, uint64_t index) {
  return rto_mix(rto_mix(seed + RTO_GOLDEN) ^ ((domain << 56) | index));
}
static inline uint32_t rto_word(uint64_t key, uint32_t c) {
  uint64_t p = rto_mix(key + (uint64_t)((c >> 1) + 1u) * RTO_GOLDEN);
  return (c & 1u) ? (uint32_t)(p >> 32) : (uint32_t)p;
}
#endif
