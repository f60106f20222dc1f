// tests/jams_stub/pcg_random.hpp — TEST INFRASTRUCTURE: jams/common.h needs a type `pcg32` (the PCG library is fetched by the reference's
// cmake at configure time and is absent here).  Compile-only stand-in with the interface common.h touches.
#ifndef JB_STUB_PCG_RANDOM_HPP
#define JB_STUB_PCG_RANDOM_HPP
#include <cstdint>
#include <iosfwd>
#include <random>
class pcg32 {
 public:
  using result_type = std::uint32_t;
  pcg32() = default;
  template <class SeedSeq> explicit pcg32(SeedSeq &&) {}
  static constexpr result_type min() { return 0u; }
  static constexpr result_type max() { return 0xffffffffu; }
  result_type operator()();
};
std::ostream &operator<<(std::ostream &, const pcg32 &);
std::istream &operator>>(std::istream &, pcg32 &);
#endif
