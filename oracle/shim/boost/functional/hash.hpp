// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for <boost/functional/hash.hpp> (Boost.ContainerHash) covering what the
// reference instantiates: boost::hash<int>, boost::hash<dynamic_bitset<>> (through
// hash_value found by ADL), hash_range over std::vector<unsigned long>, hash_combine.
//
// The reference does not pin a Boost version and the 64-bit hash_combine changed in 1.81,
// so both published algorithms are restated and selected at run time:
//   variant 171 (Boost 1.71 .. 1.80): k*=m; k^=k>>47; k*=m; h^=k; h*=m; h+=0xe6546b64
//                                     with m = 0xc6a4a7935bd1e995
//   variant 181 (Boost >= 1.81):      h = mix(h + 0x9e3779b9 + k),
//                                     mix(x): x^=x>>32; x*=M; x^=x>>32; x*=M; x^=x>>28
//                                     with M = 0x0e9846af9b1a615d
// Integers hash to themselves in every version.
#ifndef ORACLE_SHIM_FUNCTIONAL_HASH_HPP
#define ORACLE_SHIM_FUNCTIONAL_HASH_HPP

#include <cstddef>
#include <cstdint>
#include <vector>

namespace boost {

namespace shim {
inline int& hash_variant() {
  static int v = 181;
  return v;
}
inline std::uint64_t combine171(std::uint64_t h, std::uint64_t k) {
  const std::uint64_t m = 0xc6a4a7935bd1e995ULL;
  k *= m; k ^= k >> 47; k *= m;
  h ^= k; h *= m; h += 0xe6546b64ULL;
  return h;
}
inline std::uint64_t combine181(std::uint64_t h, std::uint64_t k) {
  const std::uint64_t M = 0x0e9846af9b1a615dULL;
  std::uint64_t x = h + 0x9e3779b9ULL + k;
  x ^= x >> 32; x *= M; x ^= x >> 32; x *= M; x ^= x >> 28;
  return x;
}
}  // namespace shim

inline std::size_t hash_value(int v) { return static_cast<std::size_t>(v); }
inline std::size_t hash_value(unsigned int v) { return static_cast<std::size_t>(v); }
inline std::size_t hash_value(long v) { return static_cast<std::size_t>(v); }
inline std::size_t hash_value(unsigned long v) { return static_cast<std::size_t>(v); }
inline std::size_t hash_value(unsigned long long v) { return static_cast<std::size_t>(v); }

template <class T> struct hash;

template <class T>
inline void hash_combine(std::size_t& seed, const T& v);

template <class It>
inline void hash_range(std::size_t& seed, It first, It last) {
  for (; first != last; ++first) hash_combine(seed, *first);
}
template <class It>
inline std::size_t hash_range(It first, It last) {
  std::size_t seed = 0;
  hash_range(seed, first, last);
  return seed;
}
template <class T, class A>
inline std::size_t hash_value(const std::vector<T, A>& v) {
  return hash_range(v.begin(), v.end());
}

template <class T>
struct hash {
  std::size_t operator()(const T& v) const { return hash_value(v); }  // ADL for class types
};

template <class T>
inline void hash_combine(std::size_t& seed, const T& v) {
  const std::size_t k = boost::hash<T>()(v);
  seed = (shim::hash_variant() == 171) ? shim::combine171(seed, k) : shim::combine181(seed, k);
}

}  // namespace boost

#endif
