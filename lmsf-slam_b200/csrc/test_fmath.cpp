// test_fmath.cpp — host check of fmath.cuh against the C library's atan2f (test infrastructure; run by
// tests/test_abi.py, no GPU): every bit of every result must agree.  Build: g++ -O2 -ffp-contract=off -x c++ test_fmath.cpp.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fmath.cuh"

static uint64_t s = 0x9E3779B97F4A7C15ull;
static inline uint64_t rng() {
  s ^= s << 13;
  s ^= s >> 7;
  s ^= s << 17;
  return s;
}
static inline float uni(float lo, float hi) { return lo + (hi - lo) * (float)((rng() >> 40) * (1.0 / 16777216.0)); }

int main(int argc, char** argv) {
  long n = argc > 1 ? atol(argv[1]) : 20000000L;
  long bad = 0, total = 0;
  auto check = [&](float y, float x) {
    float a = std::atan2(y, x), b = lm::atan2f_fdlibm(y, x);
    uint32_t ia, ib;
    memcpy(&ia, &a, 4);
    memcpy(&ib, &b, 4);
    ++total;
    if (ia != ib && !(a != a && b != b)) {
      if (bad < 10) printf("atan2f(%a, %a): libc %a  port %a\n", y, x, a, b);
      ++bad;
    }
  };
  const float sp[] = {0.0f, -0.0f, 1.0f, -1.0f, INFINITY, -INFINITY, NAN, 1e-45f, -1e-45f, 1e-30f, 3.4e38f, -3.4e38f,
                      0.4375f, 0.6875f, 1.1875f, 2.4375f, 33554432.0f, 1.8626451e-9f, 0.5f, 1.5f, 2.0f};
  for (float y : sp)
    for (float x : sp) check(y, x);
  for (long i = 0; i < n; ++i) {
    check(uni(-100.f, 100.f), uni(-100.f, 100.f));                 // lidar-like coordinates, all quadrants
    uint32_t by = (uint32_t)rng(), bx = (uint32_t)rng();           // arbitrary bit patterns
    float y, x;
    memcpy(&y, &by, 4);
    memcpy(&x, &bx, 4);
    check(y, x);
    float r = uni(0.f, 4.f);                                       // ratios across the reduction breakpoints
    float d = uni(1.f, 80.f);
    check(r * d, (rng() & 1) ? d : -d);
  }
  printf("arguments %ld  bit mismatches %ld\n", total, bad);
  return bad ? 1 : 0;
}
