// GPU-side ingest of Shifu's normalised training text: what the reference does in
// interpreted Python, one float() call per cell (load_data, res/ssgd_monitor.py:348-454):
//   line.split('|') -> selected columns -> float(cell) ; target float(columns[target]) ; weight column with
//   "negative -> 1.0, absent -> 1.0".
// Here: one thread per line walks its bytes in 16-byte register chunks through a small state machine and converts
// every selected cell with Clinger's exact fast path (mantissa < 2^53, |decimal exponent| <= 22: ONE correctly rounded
// double multiply / divide, then the same double -> float cast numpy applies when the reference feeds fp32
// placeholders).  Cells outside the fast path (more than 15-19 significant digits, huge exponents, "nan"/"inf",
// anything float() would have to think about) are NOT guessed: they are appended to a small list and re-parsed by the
// caller with the reference's own float().  HBM-bound byte work: no tensor cores, coalescing comes from consecutive
// 16-byte chunks per thread (every 32-byte sector fetched is fully consumed).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD inline
#endif

namespace sb {

// column roles in col_map: SB_COL_SKIP / SB_COL_TARGET / SB_COL_WEIGHT come from include/shifu_b200.h

struct CellFlag {      // a cell the fast path declined (or a malformed line); resolved by the caller
  int64_t row;
  int32_t slot;        // feature index >= 0, SB_COL_TARGET, SB_COL_WEIGHT; -100: line has too few / too many features
  int32_t len;
  int64_t offset;      // byte offset of the cell in the text
};

struct NumState {
  uint64_t mant;       // decimal significand digits accumulated (up to 19)
  int32_t digits;      // significant digits consumed into mant
  int32_t dropped;     // integer-part digits beyond 19 (scale up), only counted
  int32_t frac;        // digits after the decimal point that went into mant
  int32_t exp;         // explicit exponent value
  int8_t phase;        // 0 leading ws, 1 sign seen, 2 int digits, 3 frac digits, 4 'e' seen, 5 exp sign, 6 exp digits, 7 trailing ws
  int8_t neg, exp_neg, bad, any_digit, inexact;
};

SB_HD void num_reset(NumState& s) {
  s.mant = 0; s.digits = 0; s.dropped = 0; s.frac = 0; s.exp = 0; s.phase = 0; s.neg = 0; s.exp_neg = 0; s.bad = 0;
  s.any_digit = 0; s.inexact = 0;
}

SB_HD bool is_ws(unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// feed one byte of a cell (delimiters excluded)
SB_HD void num_feed(NumState& s, unsigned char c) {
  if (s.bad) return;
  if (c >= '0' && c <= '9') {
    const int d = c - '0';
    if (s.phase <= 2) {
      s.phase = 2; s.any_digit = 1;
      if (s.mant == 0 && d == 0) return;                       // leading zeros
      if (s.digits < 19) { s.mant = s.mant * 10 + d; ++s.digits; }
      else { ++s.dropped; if (d) s.inexact = 1; }
    } else if (s.phase == 3) {
      s.any_digit = 1;
      if (s.mant == 0 && d == 0) { ++s.frac; return; }        // 0.000ddd: zeros only shift the exponent
      if (s.digits < 19) { s.mant = s.mant * 10 + d; ++s.digits; ++s.frac; }
      else if (d) s.inexact = 1;
    } else if (s.phase >= 4 && s.phase <= 6) {
      s.phase = 6;
      if (s.exp < 100000) s.exp = s.exp * 10 + d;
    } else s.bad = 1;
    return;
  }
  if (is_ws(c)) {
    if (s.phase == 0 || s.phase == 7) return;
    if (s.phase == 1 || s.phase == 4 || s.phase == 5) { s.bad = 1; return; }
    s.phase = 7;
    return;
  }
  if (c == '+' || c == '-') {
    if (s.phase == 0) { s.phase = 1; s.neg = (c == '-'); }
    else if (s.phase == 4) { s.phase = 5; s.exp_neg = (c == '-'); }
    else s.bad = 1;
    return;
  }
  if (c == '.') {
    if (s.phase <= 2) s.phase = 3; else s.bad = 1;
    return;
  }
  if (c == 'e' || c == 'E') {
    if ((s.phase == 2 || s.phase == 3) && s.any_digit) s.phase = 4; else s.bad = 1;
    return;
  }
  s.bad = 1;   // nan, inf, underscores, hex ... -> caller's float()
}

// exact powers of ten representable in double
SB_HD double pow10_exact(int k) {
  const double t[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17,
                        1e18, 1e19, 1e20, 1e21, 1e22};
  return t[k];
}

// finish a cell: returns true and the correctly rounded float(double(text)) when the fast path applies
SB_HD bool num_finish(const NumState& s, float* out) {
  if (s.bad || !s.any_digit || s.phase == 1 || s.phase == 4 || s.phase == 5) return false;
  if (s.inexact) return false;
  if (s.mant == 0) { *out = s.neg ? -0.0f : 0.0f; return true; }
  if (s.mant > (1ull << 53)) return false;
  const long long e10 = static_cast<long long>(s.exp_neg ? -s.exp : s.exp) - s.frac + s.dropped;
  if (e10 < -22 || e10 > 22) return false;
  double v = static_cast<double>(s.mant);                       // exact (mant <= 2^53)
  v = (e10 >= 0) ? v * pow10_exact(static_cast<int>(e10)) : v / pow10_exact(static_cast<int>(-e10));  // one rounding
  if (s.neg) v = -v;
  *out = static_cast<float>(v);                                 // the numpy float64 -> float32 cast of the feed
  return true;
}

}  // namespace sb
