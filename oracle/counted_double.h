// TEST INFRASTRUCTURE — a double that counts the arithmetic performed on it.  oracle/pose_oracle_counted.cpp compiles the CPU
// restatement (pose_oracle.cpp) a second time with `double` spelled `Counted`, which turns the oracle into an exact FP64
// operation counter for bench.py's K2/K3 rooflines (SURVEY.md §8d asks for counts "from an instrumented build of the CPU
// restatement").  Classes: add (+, -, negation is free), mul, div, sqrt, cmp (comparisons), and `special` = calls of
// transcendental library functions (exp, log, sin, cos, atan2, acos, pow with a non-integer exponent, cbrt, hypot).
// std::complex<Counted> instantiates libstdc++'s generic complex algorithms, so complex pow / sqrt / division are counted by
// their real-arithmetic decomposition there (glibc's csqrt / libgcc's __divdc3, which std::complex<double> calls, perform a
// few operations more or less; the difference is below 2 % of a P3P solve).
#pragma once
#include <cmath>
#include <complex>
#include <limits>

struct OpCounts { unsigned long long add, mul, div, sqrt, cmp, special; };
inline OpCounts& op_counts() { static OpCounts c = {0, 0, 0, 0, 0, 0}; return c; }

struct Counted {
  double v;
  Counted() : v(0) {}
  Counted(double x) : v(x) {}
  Counted(float x) : v(x) {}
  Counted(int x) : v(x) {}
  Counted(unsigned x) : v(x) {}
  Counted(long x) : v((double)x) {}
  Counted(unsigned long x) : v((double)x) {}
  explicit operator int() const { return (int)v; }
  explicit operator unsigned() const { return (unsigned)v; }
  explicit operator float() const { return (float)v; }
  explicit operator bool() const { return v != 0; }
  double raw() const { return v; }
  Counted operator-() const { return Counted(-v); }
  Counted& operator+=(Counted o) { ++op_counts().add; v += o.v; return *this; }
  Counted& operator-=(Counted o) { ++op_counts().add; v -= o.v; return *this; }
  Counted& operator*=(Counted o) { ++op_counts().mul; v *= o.v; return *this; }
  Counted& operator/=(Counted o) { ++op_counts().div; v /= o.v; return *this; }
};
#define CNT_BIN(op, field) \
  inline Counted operator op(Counted a, Counted b) { ++op_counts().field; return Counted(a.v op b.v); } \
  inline Counted operator op(Counted a, double b) { ++op_counts().field; return Counted(a.v op b); } \
  inline Counted operator op(double a, Counted b) { ++op_counts().field; return Counted(a op b.v); } \
  inline Counted operator op(Counted a, int b) { ++op_counts().field; return Counted(a.v op b); } \
  inline Counted operator op(int a, Counted b) { ++op_counts().field; return Counted(a op b.v); } \
  inline Counted operator op(Counted a, unsigned b) { ++op_counts().field; return Counted(a.v op b); } \
  inline Counted operator op(Counted a, unsigned long b) { ++op_counts().field; return Counted(a.v op (double)b); }
CNT_BIN(+, add) CNT_BIN(-, add) CNT_BIN(*, mul) CNT_BIN(/, div)
#undef CNT_BIN
#define CNT_CMP(op) \
  inline bool operator op(Counted a, Counted b) { ++op_counts().cmp; return a.v op b.v; } \
  inline bool operator op(Counted a, double b) { ++op_counts().cmp; return a.v op b; } \
  inline bool operator op(double a, Counted b) { ++op_counts().cmp; return a op b.v; } \
  inline bool operator op(Counted a, int b) { ++op_counts().cmp; return a.v op b; } \
  inline bool operator op(int a, Counted b) { ++op_counts().cmp; return a op b.v; }
CNT_CMP(<) CNT_CMP(>) CNT_CMP(<=) CNT_CMP(>=) CNT_CMP(==) CNT_CMP(!=)
#undef CNT_CMP

#define CNT_FN1(name, field) inline Counted name(Counted a) { ++op_counts().field; return Counted(::name(a.v)); }
inline Counted sqrt(Counted a) { ++op_counts().sqrt; return Counted(::sqrt(a.v)); }
inline Counted fabs(Counted a) { return Counted(::fabs(a.v)); }          // sign-bit operation: free
inline Counted abs(Counted a) { return Counted(::fabs(a.v)); }
CNT_FN1(exp, special) CNT_FN1(log, special) CNT_FN1(sin, special) CNT_FN1(cos, special) CNT_FN1(acos, special) CNT_FN1(cbrt, special)
CNT_FN1(tan, special) CNT_FN1(atan, special) CNT_FN1(sinh, special) CNT_FN1(cosh, special) CNT_FN1(tanh, special)
#undef CNT_FN1
inline Counted atan2(Counted a, Counted b) { ++op_counts().special; return Counted(::atan2(a.v, b.v)); }
inline Counted hypot(Counted a, Counted b) { ++op_counts().special; return Counted(::hypot(a.v, b.v)); }
inline Counted pow(Counted a, Counted b) {      // pow(x, 2) is a multiplication in every libm; anything else is a transcendental call
  if (b.v == 2.0) { ++op_counts().mul; return Counted(a.v * a.v); }
  ++op_counts().special; return Counted(::pow(a.v, b.v));
}
inline Counted pow(Counted a, int b) { return pow(a, Counted((double)b)); }
inline Counted pow(Counted a, double b) { return pow(a, Counted(b)); }
inline bool isnan(Counted a) { return a.v != a.v; }
inline bool isinf(Counted a) { return std::isinf(a.v); }
inline bool isfinite(Counted a) { return std::isfinite(a.v); }
inline Counted fmin(Counted a, Counted b) { ++op_counts().cmp; return Counted(::fmin(a.v, b.v)); }
inline Counted fmax(Counted a, Counted b) { ++op_counts().cmp; return Counted(::fmax(a.v, b.v)); }
inline Counted copysign(Counted a, Counted b) { return Counted(::copysign(a.v, b.v)); }

namespace std {
template <> struct numeric_limits<Counted> {
  static constexpr bool is_specialized = true;
  static Counted infinity() { return Counted(numeric_limits<double>::infinity()); }
  static Counted quiet_NaN() { return Counted(numeric_limits<double>::quiet_NaN()); }
  static Counted epsilon() { return Counted(numeric_limits<double>::epsilon()); }
  static Counted min() { return Counted(numeric_limits<double>::min()); }
  static Counted max() { return Counted(numeric_limits<double>::max()); }
};
using ::sqrt; using ::fabs; using ::abs; using ::exp; using ::log; using ::sin; using ::cos; using ::acos; using ::atan2; using ::pow;
using ::hypot; using ::isnan; using ::isinf; using ::isfinite; using ::cbrt; using ::tan; using ::atan; using ::sinh; using ::cosh; using ::tanh;
}

// literals next to std::complex<Counted> (template deduction would otherwise see two different scalar types)
typedef std::complex<Counted> CComplex;
inline CComplex operator/(const CComplex& a, double b) { return a / Counted(b); }
inline CComplex operator*(const CComplex& a, double b) { return a * Counted(b); }
inline CComplex operator*(double a, const CComplex& b) { return Counted(a) * b; }
inline CComplex operator+(const CComplex& a, double b) { return a + Counted(b); }
inline CComplex operator+(double a, const CComplex& b) { return Counted(a) + b; }
inline CComplex operator-(const CComplex& a, double b) { return a - Counted(b); }
inline CComplex operator-(double a, const CComplex& b) { return Counted(a) - b; }
inline CComplex operator/(double a, const CComplex& b) { return Counted(a) / b; }
namespace std {
inline CComplex pow(const CComplex& a, double b) { return std::pow(a, Counted(b)); }
inline Counted min(Counted a, double b) { ++op_counts().cmp; return Counted(a.v < b ? a.v : b); }
inline Counted min(double a, Counted b) { ++op_counts().cmp; return Counted(b.v < a ? b.v : a); }
inline Counted max(Counted a, double b) { ++op_counts().cmp; return Counted(a.v < b ? b : a.v); }
inline Counted max(double a, Counted b) { ++op_counts().cmp; return Counted(a < b.v ? b.v : a); }
}
