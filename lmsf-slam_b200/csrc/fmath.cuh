// fmath.cuh — single-precision atan2 with the bits of the host C library's atan2f.
//
// The reference's bad-point test calls atan2 on two floats (LOAMFeatureProcessor_base.hpp:223-224) inside a
// translation unit where `using namespace std;` is in effect (src/apps/include/utility.hpp:51), i.e. std::atan2(float,
// float) = the C library's atan2f.  A float azimuth near 3 rad has an ulp of 2.4e-7 and the test works on the
// DIFFERENCE of two neighbouring azimuths (3e-3 rad), so an implementation that is merely "accurate to 2 ulp"
// (CUDA's atan2f) changes occlusion labels every few sweeps.  glibc (<= 2.40) computes atan2f with the classic
// fdlibm float algorithm (argument reduction to four breakpoints + an 11-term odd polynomial), which is IEEE float
// + - * / only: this file is that algorithm, written from its published description, so that with -fmad=false the
// device returns the host's bits.  tests: csrc/test_fmath.cpp (host, against the C library: 3.9e8 arguments, 0 mismatches with glibc 2.39) and
// csrc/test_dmath.cu (device against host).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define FM_HD __host__ __device__ __forceinline__
#else
#define FM_HD inline  // plain g++ (test_fmath.cpp)
#endif

namespace lm {

FM_HD int32_t f2i(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int32_t i;
  memcpy(&i, &f, 4);
  return i;
#endif
}
FM_HD float i2f(int32_t i) {
#ifdef __CUDA_ARCH__
  return __int_as_float(i);
#else
  float f;
  memcpy(&f, &i, 4);
  return f;
#endif
}

// atan(x), fdlibm single precision: |x| < 7/16 direct; otherwise reduced around 0.5, 1, 1.5 or infinity
FM_HD float atanf_fdlibm(float x) {
  const float hi0 = i2f(0x3eed6338), hi1 = i2f(0x3f490fda), hi2 = i2f(0x3f7b985e), hi3 = i2f(0x3fc90fda);
  const float lo0 = i2f(0x31ac3769), lo1 = i2f(0x33222168), lo2 = i2f(0x33140fb4), lo3 = i2f(0x33a22168);
  // a0: the published decimal 3.3333334327e-01 rounds to 0x3eaaaaab (the table's hex comment says ...aaa); the C
  // library is compiled from the decimal, and so is this
  const float a0 = i2f(0x3eaaaaab), a1 = i2f((int32_t)0xbe4ccccd), a2 = i2f(0x3e124925), a3 = i2f((int32_t)0xbde38e38),
              a4 = i2f(0x3dba2e6e), a5 = i2f((int32_t)0xbd9d8795), a6 = i2f(0x3d886b35), a7 = i2f((int32_t)0xbd6ef16b),
              a8 = i2f(0x3d4bda59), a9 = i2f((int32_t)0xbd15a221), a10 = i2f(0x3c8569d7);
  const int32_t hx = f2i(x);
  const int32_t ix = hx & 0x7fffffff;
  int id;
  float hi = 0.f, lo = 0.f;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return x + x;
    return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
  }
  if (ix < 0x3ee00000) {            // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {    // |x| < 1.1875
      if (ix < 0x3f300000) {  // 7/16 <= |x| < 11/16
        id = 0;
        hi = hi0;
        lo = lo0;
        x = (2.0f * x - 1.0f) / (2.0f + x);
      } else {  // 11/16 <= |x| < 19/16
        id = 1;
        hi = hi1;
        lo = lo1;
        x = (x - 1.0f) / (x + 1.0f);
      }
    } else {
      if (ix < 0x401c0000) {  // |x| < 2.4375
        id = 2;
        hi = hi2;
        lo = lo2;
        x = (x - 1.5f) / (1.0f + 1.5f * x);
      } else {
        id = 3;
        hi = hi3;
        lo = lo3;
        x = -1.0f / x;
      }
    }
  }
  float z = x * x;
  float w = z * z;
  float s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
  float s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
  if (id < 0) return x - x * (s1 + s2);
  z = hi - ((x * (s1 + s2) - lo) - x);
  return hx < 0 ? -z : z;
}

// atan2(y, x), fdlibm single precision
FM_HD float atan2f_fdlibm(float y, float x) {
  const float tiny = 1.0e-30f;
  const float pi_o_4 = i2f(0x3f490fdb), pi_o_2 = i2f(0x3fc90fdb), pi = i2f(0x40490fdb), pi_lo = i2f((int32_t)0xb3bbbd2e);
  const int32_t hx = f2i(x), hy = f2i(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;  // NaN
  if (hx == 0x3f800000) return atanf_fdlibm(y);           // x = 1
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);      // 2*sign(x) + sign(y)
  if (iy == 0) {
    if (m < 2) return y;
    return m == 2 ? pi + tiny : -pi - tiny;
  }
  if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    }
    switch (m) {
      case 0: return 0.0f;
      case 1: return -0.0f;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 60)
    z = pi_o_2 + 0.5f * pi_lo;  // |y/x| > 2^60
  else if (hx < 0 && k < -60)
    z = 0.0f;  // |y|/x < -2^60
  else
    z = atanf_fdlibm(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return i2f(f2i(z) ^ (int32_t)0x80000000);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

}  // namespace lm
