// fastlog.cuh -- natural logarithm for the barrier kernels (barrier.cu).
//
// The barrier energy b(d) = -(d - m)^2 log(d / m) (Energy_admm.h:46-96) and its derivatives (Gradient_admm.h:331-407) are
// evaluated for every (control point, plane) term inside the band: ncu put 47 % of the stall samples of k_row_energy and 38 %
// of k_row_grad on the logarithm (profiles/r02_hot_batch1024_*).  The library log() is a chain of 29 dependent FP64
// operations plus special-case handling; the arguments here are always positive normal numbers (0 < d/m <= 1), so:
//
//   x = 2^e * f, f in [c, 2c), c = 0.70834 (high word 0x3fe6aaab: puts 1.0 two thirds into its bin, i.e. in the middle of
//   the bin's VALUE range)                           (integer arithmetic on the high word, as the library does)
//   f falls in one of 128 bins of the high word; bin i has (inv_i, T_i = -log(inv_i))         (csrc/logtab.inc)
//   r = f * inv_i - 1   (one FMA, exact product), |r| <= 2^-8
//   log x = (e * ln2_hi + T_i) + (e * ln2_lo + r + r^2 * (-1/2 + r/3 - r^2/4 + r^3/5 - r^4/6 + r^5/7))
//
// 12 FP64 operations, dependency depth 6.  The bin that contains 1 has inv = 1, T = 0: next to x = 1 the result is the
// series in r = x - 1 alone and keeps its relative accuracy.  Truncation r^7/8 <= 2e-18 relative; measured against the
// host's long-double logarithm: <= 1 ulp over the band (tests/test_cpu_checks.py::test_fast_log).  Anything that is not a
// positive normal number goes to the library function.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TOB_FL_HD __host__ __device__ __forceinline__
#else
#define TOB_FL_HD inline
#endif

namespace tob {

struct LogTabEntry { double inv, t; };
#define TOB_LOGTAB_N 128
static const LogTabEntry kLogTabHost[TOB_LOGTAB_N] = {
#include "logtab.inc"
};

// tab: TOB_LOGTAB_N entries (shared memory on the device)
TOB_FL_HD double tob_log_pos(double x, const LogTabEntry* tab) {
  int hi, lo;
#if defined(__CUDA_ARCH__)
  hi = __double2hiint(x); lo = __double2loint(x);
#else
  { uint64_t u; memcpy(&u, &x, 8); hi = (int)(u >> 32); lo = (int)(u & 0xffffffffu); }
#endif
  if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log(x);   // zero, subnormal, negative, inf, NaN: never in the band
  const int t = hi - 0x3fe6aaab;
  const int e = t >> 20, mant = t & 0x000fffff;
  const int hf = mant + 0x3fe6aaab;
  double f;
#if defined(__CUDA_ARCH__)
  f = __hiloint2double(hf, lo);
#else
  { uint64_t u = ((uint64_t)(uint32_t)hf << 32) | (uint32_t)lo; memcpy(&f, &u, 8); }
#endif
  const LogTabEntry en = tab[mant >> 13];
  const double ed = (double)e;
  const double r = fma(f, en.inv, -1.0);
  const double r2 = r * r;
  const double a = fma(r, 1.0 / 3.0, -0.5), b = fma(r, 0.2, -0.25), c = fma(r, 1.0 / 7.0, -1.0 / 6.0);
  const double q = fma(r2, fma(r2, c, b), a);
  const double hi_part = fma(ed, 0x1.62e42feep-1, en.t);                    // ln2 high part (32 significant bits: e * it is exact)
  const double lo_part = fma(ed, 0x1.a39ef35793c76p-33, fma(r2, q, r));     // ln2 - high part
  return hi_part + lo_part;
}

}  // namespace tob
