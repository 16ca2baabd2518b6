// Device-side Kneip P3P + Ferrari quartic, FP64.
//
// What it computes is P3P::computePoses / P3P::solveQuartic
// (/root/reference/monocular_pose_estimator_lib/src/p3p.cpp:65-236, :238-286).  The reference evaluates the
// quartic with std::complex<double>; there is no std::complex on the device, so the few complex
// operations it needs (pow(z, real), sqrt, division, real scalings) are written out here following the
// semantics of the host toolchain the reference is built with:
//   pow(complex, real)  -> libstdc++ <complex>: real pow for positive-real bases, else polar(exp(y*log|z|), y*arg z)
//   sqrt(complex)       -> glibc csqrt
//   complex / complex   -> libgcc __divdc3 (Smith's method with NaN recovery)
// Compiled with -fmad=false so that every + and * rounds once, like the reference's -O3 x86-64 build
// without -march=native (L/CMakeLists.txt:5-6).
// ---------------------------------------------------------------------------------------------------------------------------
// The P3P parametrisation and the quartic coefficient expressions restated in this file follow Laurent Kneip's algorithm as
// distributed with the reference (monocular_pose_estimator_lib/src/p3p.cpp), whose licence requires this notice to be retained:
//
//   Copyright (c) 2011, Laurent Kneip, ETH Zurich.  All rights reserved.
//
//   Redistribution and use in source and binary forms, with or without modification, are permitted provided that the following
//   conditions are met:
//     * Redistributions of source code must retain the above copyright notice, this list of conditions and the following
//       disclaimer.
//     * Redistributions in binary form must reproduce the above copyright notice, this list of conditions and the following
//       disclaimer in the documentation and/or other materials provided with the distribution.
//     * Neither the name of ETH Zurich nor the names of its contributors may be used to endorse or promote products derived
//       from this software without specific prior written permission.
//
//   THIS SOFTWARE IS PROVIDED BY THE COPYRIGHT HOLDERS AND CONTRIBUTORS "AS IS" AND ANY EXPRESS OR IMPLIED WARRANTIES, INCLUDING,
//   BUT NOT LIMITED TO, THE IMPLIED WARRANTIES OF MERCHANTABILITY AND FITNESS FOR A PARTICULAR PURPOSE ARE DISCLAIMED.  IN NO
//   EVENT SHALL ETH ZURICH BE LIABLE FOR ANY DIRECT, INDIRECT, INCIDENTAL, SPECIAL, EXEMPLARY, OR CONSEQUENTIAL DAMAGES
//   (INCLUDING, BUT NOT LIMITED TO, PROCUREMENT OF SUBSTITUTE GOODS OR SERVICES; LOSS OF USE, DATA, OR PROFITS; OR BUSINESS
//   INTERRUPTION) HOWEVER CAUSED AND ON ANY THEORY OF LIABILITY, WHETHER IN CONTRACT, STRICT LIABILITY, OR TORT (INCLUDING
//   NEGLIGENCE OR OTHERWISE) ARISING IN ANY WAY OUT OF THE USE OF THIS SOFTWARE, EVEN IF ADVISED OF THE POSSIBILITY OF SUCH DAMAGE.
// ---------------------------------------------------------------------------------------------------------------------------
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mpe {

struct cplx { double re, im; };

__device__ __forceinline__ cplx c_make(double re, double im) { cplx z; z.re = re; z.im = im; return z; }
__device__ __forceinline__ cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx c_neg(cplx a) { return c_make(-a.re, -a.im); }
__device__ __forceinline__ cplx c_scale(double s, cplx a) { return c_make(s * a.re, s * a.im); }   // real * complex
__device__ __forceinline__ cplx c_divr(cplx a, double s) { return c_make(a.re / s, a.im / s); }   // complex / real
__device__ __forceinline__ cplx r_add(double r, cplx a) { return c_make(r + a.re, a.im); }        // real + complex
__device__ __forceinline__ cplx r_sub(double r, cplx a) { return c_make(r - a.re, -a.im); }       // real - complex

// std::polar(rho, theta)
__device__ __forceinline__ cplx c_polar(double rho, double theta) {
  double s, c;
  sincos(theta, &s, &c);
  return c_make(rho * c, rho * s);
}

// std::pow(const complex<double>&, const double&) as libstdc++ defines it, specialised for the three exponents
// solveQuartic uses (2.0, 3.0, 1.0/3.0):
//   positive-real base -> real pow: x*x, x*x*x (what a correctly rounded pow returns for x^2; within 1 ulp for x^3) and cbrt;
//   otherwise          -> polar(exp(y * log|z|), y * arg z).  For a negative-real base arg z = +-pi, so the angle y*pi is a
//                         constant: its sine/cosine are the values glibc's sincos returns for fl(y*pi) (hex literals below),
//                         and log|z| needs no hypot.
template <int kExp>   // 2, 3, or 0 for 1/3
__device__ __forceinline__ cplx c_pow_real(cplx z) {
  const double y = (kExp == 2) ? 2.0 : (kExp == 3) ? 3.0 : (1.0 / 3.0);
  if (z.im == 0.0 && z.re > 0.0) {
    double r = (kExp == 2) ? z.re * z.re : (kExp == 3) ? z.re * z.re * z.re : cbrt(z.re);
    return c_make(r, 0.0);
  }
  if (z.im == 0.0 && z.re < 0.0) {
    // sin/cos of fl(y * pi) from glibc 2.39 (= what the reference's libm returns); sign of the angle follows arg z = copysign(pi, im)
    const double sn = (kExp == 2) ? -0x1.1a62633145c07p-52 : (kExp == 3) ? 0x1.a79394c9e8a0ap-52 : 0x1.bb67ae8584caap-1;
    const double cs = (kExp == 2) ? 1.0 : (kExp == 3) ? -1.0 : 0x1.0000000000001p-1;
    double rho = exp(y * log(-z.re));
    return c_make(rho * cs, rho * (signbit(z.im) ? -sn : sn));
  }
  double lr, li;
  if (z.re == 0.0 && z.im == 0.0) {          // glibc clog(0): (-inf, signbit(re) ? pi : 0)
    lr = -1.0 / fabs(z.re);
    li = signbit(z.re) ? 3.14159265358979323846 : 0.0;
    li = copysign(li, z.im);
  } else {
    lr = log(hypot(z.re, z.im));
    li = atan2(z.im, z.re);
  }
  return c_polar(exp(y * lr), y * li);
}

// glibc csqrt (finite, non-extreme magnitudes; Inf/NaN classes handled as glibc does)
__device__ __forceinline__ cplx c_sqrt(cplx z) {
  double re = z.re, im = z.im;
  bool re_nan = isnan(re), im_nan = isnan(im), re_inf = isinf(re), im_inf = isinf(im);
  if (re_nan || im_nan || re_inf || im_inf) {
    if (im_inf) return c_make(HUGE_VAL, im);
    if (re_inf) {
      if (re < 0.0) return c_make(im_nan ? nan("") : 0.0, copysign(HUGE_VAL, im));
      return c_make(re, im_nan ? nan("") : copysign(0.0, im));
    }
    return c_make(nan(""), nan(""));
  }
  if (im == 0.0) {
    if (re < 0.0) return c_make(0.0, copysign(sqrt(-re), im));
    return c_make(fabs(sqrt(re)), copysign(0.0, im));
  }
  if (re == 0.0) {
    double r = sqrt(0.5 * fabs(im));
    return c_make(r, copysign(r, im));
  }
  double d = hypot(re, im), r, s;
  if (re > 0.0) {
    r = sqrt(0.5 * (d + re));
    s = 0.5 * (im / r);
  } else {
    s = sqrt(0.5 * (d - re));
    r = fabs(0.5 * (im / s));
  }
  return c_make(r, copysign(s, im));
}

// libgcc __divdc3: (a + ib) / (c + id)
__device__ __forceinline__ cplx c_div(cplx num, cplx den) {
  double a = num.re, b = num.im, c = den.re, d = den.im;
  double ratio, denom, x, y;
  if (fabs(c) < fabs(d)) {
    ratio = c / d;
    denom = (c * ratio) + d;
    x = ((a * ratio) + b) / denom;
    y = ((b * ratio) - a) / denom;
  } else {
    ratio = d / c;
    denom = (d * ratio) + c;
    x = ((b * ratio) + a) / denom;
    y = (b - (a * ratio)) / denom;
  }
  if (isnan(x) && isnan(y)) {
    if (c == 0.0 && d == 0.0 && (!isnan(a) || !isnan(b))) {
      x = copysign(HUGE_VAL, c) * a;
      y = copysign(HUGE_VAL, c) * b;
    } else if ((isinf(a) || isinf(b)) && isfinite(c) && isfinite(d)) {
      a = copysign(isinf(a) ? 1.0 : 0.0, a);
      b = copysign(isinf(b) ? 1.0 : 0.0, b);
      x = HUGE_VAL * (a * c + b * d);
      y = HUGE_VAL * (b * c - a * d);
    } else if ((isinf(c) || isinf(d)) && isfinite(a) && isfinite(b)) {
      c = copysign(isinf(c) ? 1.0 : 0.0, c);
      d = copysign(isinf(d) ? 1.0 : 0.0, d);
      x = 0.0 * (a * c + b * d);
      y = 0.0 * (b * c - a * d);
    }
  }
  return c_make(x, y);
}

// p3p.cpp:238-286
__device__ __forceinline__ void solve_quartic(const double factors[5], double real_roots[4]) {
  double A = factors[0], B = factors[1], C = factors[2], D = factors[3], E = factors[4];
  double A_pw2 = A * A, B_pw2 = B * B;
  double A_pw3 = A_pw2 * A, B_pw3 = B_pw2 * B;
  double A_pw4 = A_pw3 * A, B_pw4 = B_pw3 * B;

  double alpha = -3 * B_pw2 / (8 * A_pw2) + C / A;
  double beta = B_pw3 / (8 * A_pw3) - B * C / (2 * A_pw2) + D / A;
  double gamma = -3 * B_pw4 / (256 * A_pw4) + B_pw2 * C / (16 * A_pw3) - B * D / (4 * A_pw2) + E / A;

  double alpha_pw2 = alpha * alpha;
  double alpha_pw3 = alpha_pw2 * alpha;

  cplx P = c_make(-alpha_pw2 / 12 - gamma, 0.0);
  cplx Q = c_make(-alpha_pw3 / 108 + alpha * gamma / 3 - (beta * beta) / 8, 0.0);
  // R = -Q/2 + sqrt(pow(Q,2)/4 + pow(P,3)/27)
  cplx R = c_add(c_divr(c_neg(Q), 2.0), c_sqrt(c_add(c_divr(c_pow_real<2>(Q), 4.0), c_divr(c_pow_real<3>(P), 27.0))));
  cplx U = c_pow_real<0>(R);
  cplx y;
  double m56a = -5.0 * alpha / 6.0;
  if (U.re == 0.0)
    y = r_sub(m56a, c_pow_real<0>(Q));
  else
    y = c_add(r_sub(m56a, c_div(P, c_scale(3.0, U))), U);

  cplx w = c_sqrt(r_add(alpha, c_scale(2.0, y)));
  double mB4A = -B / (4.0 * A);
  cplx two_beta = c_make(2.0 * beta, 0.0);
  cplx q = c_div(two_beta, w);
  cplx base = r_add(3.0 * alpha, c_scale(2.0, y));          // 3a + 2y
  cplx s_plus = c_sqrt(c_neg(c_add(base, q)));               // sqrt(-(3a + 2y + 2b/w))
  cplx s_minus = c_sqrt(c_neg(c_sub(base, q)));              // sqrt(-(3a + 2y - 2b/w))
  real_roots[0] = mB4A + 0.5 * (w.re + s_plus.re);
  real_roots[1] = mB4A + 0.5 * (w.re - s_plus.re);
  real_roots[2] = mB4A + 0.5 * (-w.re + s_minus.re);
  real_roots[3] = mB4A + 0.5 * (-w.re - s_minus.re);
}

struct v3 { double x, y, z; };
__device__ __forceinline__ v3 v_make(double x, double y, double z) { v3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ v3 v_sub(v3 a, v3 b) { return v_make(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 v_add(v3 a, v3 b) { return v_make(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 v_cross(v3 a, v3 b) { return v_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double v_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double v_norm(v3 a) { return sqrt(v_dot(a, a)); }
__device__ __forceinline__ v3 v_divs(v3 a, double s) { return v_make(a.x / s, a.y / s, a.z / s); }

// Shared (root-independent) part of computePoses: everything up to the quartic (p3p.cpp:65-190).
struct P3PSetup {
  v3 e1, e2, e3;     // rows of T
  v3 n1, n2, n3;     // rows of N
  v3 P1;
  double f_1, f_2, p_1, p_2, d_12, b;
  double roots[4];
};

// rows-as-vectors matrix * vector
__device__ __forceinline__ v3 rows_mul(v3 r0, v3 r1, v3 r2, v3 v) { return v_make(v_dot(r0, v), v_dot(r1, v), v_dot(r2, v)); }

// ---- computePoses up to the quartic, in three parts so that callers can hoist what does not change between problems.
// The arithmetic (operations and their order) is that of p3p.cpp:65-190 in every composition below.

// Part W — everything that depends on the ordered world-point triple only (p3p.cpp:74-80, 124-141): the colinearity
// test, the world frame N = [n1 n2 n3]^T, P3 in that frame (p_1, p_2) and d_12.
struct P3PWorld {
  v3 n1, n2, n3, P1;
  double p_1, p_2, d_12;
  double cross_norm;      // |(P2-P1) x (P3-P1)|: 0 -> colinear (the reference's reject test)
};
__device__ __forceinline__ void p3p_world_frame(v3 P1, v3 P2, v3 P3, P3PWorld& W) {
  v3 temp1 = v_sub(P2, P1), temp2 = v_sub(P3, P1);
  W.cross_norm = v_norm(v_cross(temp1, temp2));
  v3 n1 = v_sub(P2, P1);
  n1 = v_divs(n1, v_norm(n1));
  v3 n3 = v_cross(n1, v_sub(P3, P1));
  n3 = v_divs(n3, v_norm(n3));
  v3 n2 = v_cross(n3, n1);
  v3 P3n = rows_mul(n1, n2, n3, v_sub(P3, P1));
  W.n1 = n1; W.n2 = n2; W.n3 = n3; W.P1 = P1;
  W.p_1 = P3n.x; W.p_2 = P3n.y;
  W.d_12 = v_norm(v_sub(P2, P1));
}

// Part C — everything that depends on the three bearing vectors only (p3p.cpp:88-121, 142-154): the camera frame
// T = [e1 e2 e3]^T after the possible exchange of points 1 and 2, f_1, f_2 and b.  swap != 0: the caller must exchange
// world points 1 and 2 (p3p.cpp:117-119).
struct P3PCamera {
  v3 e1, e2, e3;
  double f_1, f_2, b;
  double sin12;           // |f1 x f2| (conditioning of the frame; not part of the reference's arithmetic)
  int swap;
};
__device__ __forceinline__ void p3p_camera_frame(v3 f1, v3 f2, v3 f3in, P3PCamera& Cm) {
  v3 e1 = f1;
  v3 e3 = v_cross(f1, f2);
  double nn = v_norm(e3);
  e3 = v_divs(e3, nn);
  v3 e2 = v_cross(e3, e1);
  v3 f3 = rows_mul(e1, e2, e3, f3in);
  int swap = 0;
  if (f3.z > 0.0) {   // p3p.cpp:101-121: swap the roles of points 1 and 2
    v3 t = f1; f1 = f2; f2 = t;
    e1 = f1;
    e3 = v_cross(f1, f2);
    nn = v_norm(e3);
    e3 = v_divs(e3, nn);
    e2 = v_cross(e3, e1);
    f3 = rows_mul(e1, e2, e3, f3in);
    swap = 1;
  }
  double cos_beta = v_dot(f1, f2);
  double b = 1 / (1 - cos_beta * cos_beta) - 1;
  b = (cos_beta < 0) ? -sqrt(b) : sqrt(b);
  Cm.e1 = e1; Cm.e2 = e2; Cm.e3 = e3;
  Cm.f_1 = f3.x / f3.z;
  Cm.f_2 = f3.y / f3.z;
  Cm.b = b;
  Cm.sin12 = nn;
  Cm.swap = swap;
}

// Part Q — quartic coefficients (p3p.cpp:171-185, same association) and its roots.
__device__ __forceinline__ void p3p_quartic(double f_1, double f_2, double p_1, double p_2, double d_12, double b, double roots[4]) {
  double f_1_pw2 = f_1 * f_1, f_2_pw2 = f_2 * f_2;
  double p_1_pw2 = p_1 * p_1, p_1_pw3 = p_1_pw2 * p_1, p_1_pw4 = p_1_pw3 * p_1;
  double p_2_pw2 = p_2 * p_2, p_2_pw3 = p_2_pw2 * p_2, p_2_pw4 = p_2_pw3 * p_2;
  double d_12_pw2 = d_12 * d_12, b_pw2 = b * b;

  double factors[5];
  factors[0] = -f_2_pw2 * p_2_pw4 - p_2_pw4 * f_1_pw2 - p_2_pw4;
  factors[1] = 2 * p_2_pw3 * d_12 * b + 2 * f_2_pw2 * p_2_pw3 * d_12 * b - 2 * f_2 * p_2_pw3 * f_1 * d_12;
  factors[2] = -f_2_pw2 * p_2_pw2 * p_1_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2
      + f_2_pw2 * p_2_pw4 + p_2_pw4 * f_1_pw2 + 2 * p_1 * p_2_pw2 * d_12 + 2 * f_1 * f_2 * p_1 * p_2_pw2 * d_12 * b
      - p_2_pw2 * p_1_pw2 * f_1_pw2 + 2 * p_1 * p_2_pw2 * f_2_pw2 * d_12 - p_2_pw2 * d_12_pw2 * b_pw2
      - 2 * p_1_pw2 * p_2_pw2;
  factors[3] = 2 * p_1_pw2 * p_2 * d_12 * b + 2 * f_2 * p_2_pw3 * f_1 * d_12 - 2 * f_2_pw2 * p_2_pw3 * d_12 * b
      - 2 * p_1 * p_2 * d_12_pw2 * b;
  factors[4] = -2 * f_2 * p_2_pw2 * f_1 * p_1 * d_12 * b + f_2_pw2 * p_2_pw2 * d_12_pw2 + 2 * p_1_pw3 * d_12
      - p_1_pw2 * d_12_pw2 + f_2_pw2 * p_2_pw2 * p_1_pw2 - p_1_pw4 - 2 * f_2_pw2 * p_2_pw2 * p_1 * d_12
      + p_2_pw2 * f_1_pw2 * p_1_pw2 + f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2;
  solve_quartic(factors, roots);
}

__device__ __forceinline__ void p3p_assemble(const P3PCamera& Cm, const P3PWorld& W, P3PSetup& S) {
  S.e1 = Cm.e1; S.e2 = Cm.e2; S.e3 = Cm.e3; S.n1 = W.n1; S.n2 = W.n2; S.n3 = W.n3; S.P1 = W.P1;
  S.f_1 = Cm.f_1; S.f_2 = Cm.f_2; S.p_1 = W.p_1; S.p_2 = W.p_2; S.d_12 = W.d_12; S.b = Cm.b;
  p3p_quartic(Cm.f_1, Cm.f_2, W.p_1, W.p_2, W.d_12, Cm.b, S.roots);
}

// The whole thing for one problem.  Returns 0, or -1 if the world points are colinear (p3p.cpp:77-80).
__device__ __forceinline__ int p3p_setup(v3 f1, v3 f2, v3 f3in, v3 P1, v3 P2, v3 P3, P3PSetup& S) {
  v3 temp1 = v_sub(P2, P1), temp2 = v_sub(P3, P1);
  if (v_norm(v_cross(temp1, temp2)) == 0.0) return -1;
  P3PCamera Cm;
  p3p_camera_frame(f1, f2, f3in, Cm);
  P3PWorld W;
  if (Cm.swap) p3p_world_frame(P2, P1, P3, W); else p3p_world_frame(P1, P2, P3, W);
  p3p_assemble(Cm, W, S);
  return 0;
}

// Back-substitution of root i (p3p.cpp:193-233).  H = [R | C] row-major 3x4 (camera -> world).
// Returns false (H untouched) when a NaN in the root-dependent scalars makes every entry of R NaN, i.e. exactly in
// cases that PoseEstimator::isFinite would reject afterwards; saves the two 3x3 products for dead hypotheses.
__device__ __forceinline__ bool p3p_solution(const P3PSetup& S, int i, double H[12]) {
  double root = S.roots[i];
  double cot_alpha = (-S.f_1 * S.p_1 / S.f_2 - root * S.p_2 + S.d_12 * S.b) / (-S.f_1 * root * S.p_2 / S.f_2 + S.p_1 - S.d_12);
  double cos_theta = root;
  double sin_theta = sqrt(1 - root * root);
  double sin_alpha = sqrt(1 / (cot_alpha * cot_alpha + 1));
  double cos_alpha = sqrt(1 - sin_alpha * sin_alpha);
  if (cot_alpha < 0) cos_alpha = -cos_alpha;
  if (isnan(sin_theta) || isnan(sin_alpha)) return false;   // R(0,2) = -sin_alpha*sin_theta etc. would all be NaN

  double k = (sin_alpha * S.b + cos_alpha);
  double C0 = S.d_12 * cos_alpha * k;
  double C1 = cos_theta * S.d_12 * sin_alpha * k;
  double C2 = sin_theta * S.d_12 * sin_alpha * k;
  // C = P1 + N^T * C
  H[3] = S.P1.x + (S.n1.x * C0 + S.n2.x * C1 + S.n3.x * C2);
  H[7] = S.P1.y + (S.n1.y * C0 + S.n2.y * C1 + S.n3.y * C2);
  H[11] = S.P1.z + (S.n1.z * C0 + S.n2.z * C1 + S.n3.z * C2);

  // R (p3p.cpp:211-220); Rt = R^T
  double R00 = -cos_alpha, R01 = -sin_alpha * cos_theta, R02 = -sin_alpha * sin_theta;
  double R10 = sin_alpha, R11 = -cos_alpha * cos_theta, R12 = -cos_alpha * sin_theta;
  double R20 = 0.0, R21 = -sin_theta, R22 = cos_theta;
  // M = N^T * R^T : M(i,j) = sum_k N(k,i) * R(j,k)
  double Nt[3][3] = {{S.n1.x, S.n2.x, S.n3.x}, {S.n1.y, S.n2.y, S.n3.y}, {S.n1.z, S.n2.z, S.n3.z}};
  double Rt[3][3] = {{R00, R10, R20}, {R01, R11, R21}, {R02, R12, R22}};
  double T[3][3] = {{S.e1.x, S.e1.y, S.e1.z}, {S.e2.x, S.e2.y, S.e2.z}, {S.e3.x, S.e3.y, S.e3.z}};
  double M[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) M[r][c] = Nt[r][0] * Rt[0][c] + Nt[r][1] * Rt[1][c] + Nt[r][2] * Rt[2][c];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) H[4 * r + c] = M[r][0] * T[0][c] + M[r][1] * T[1][c] + M[r][2] * T[2][c];
  return true;
}

// PoseEstimator::isFinite on [H; 0 0 0 1] (pose_estimator.cpp:856-860): false for NaN and +-Inf.
__device__ __forceinline__ bool h_is_finite(const double H[12]) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 12; ++i) { double d = H[i] - H[i]; ok = ok && (d == d); }
  return ok;
}

// Inverse of [H; 0 0 0 1] by the general cofactor formula (the reference calls Eigen's general
// Matrix4d::inverse(), pose_estimator.cpp:486,516,660).  The terms that are multiplied by the constant
// bottom row (0 0 0 1) are dropped; the remaining products are formed in the same order as the oracle's
// full 4x4 cofactor expansion, so the result is bit-identical to it.
__device__ __forceinline__ void h_inverse(const double a[12], double inv[12]) {
  // a: 0 1 2 3 / 4 5 6 7 / 8 9 10 11 / (12 13 14 15 = 0 0 0 1)
  double i0 = a[5] * a[10] - a[9] * a[6];
  double i4 = -a[4] * a[10] + a[8] * a[6];
  double i8 = a[4] * a[9] - a[8] * a[5];
  double i1 = -a[1] * a[10] + a[9] * a[2];
  double i5 = a[0] * a[10] - a[8] * a[2];
  double i9 = -a[0] * a[9] + a[8] * a[1];
  double i2 = a[1] * a[6] - a[5] * a[2];
  double i6 = -a[0] * a[6] + a[4] * a[2];
  double i10 = a[0] * a[5] - a[4] * a[1];
  double i3 = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
  double i7 = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
  double i11 = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
  double det = a[0] * i0 + a[1] * i4 + a[2] * i8;
  double idet = 1.0 / det;
  inv[0] = i0 * idet; inv[1] = i1 * idet; inv[2] = i2 * idet; inv[3] = i3 * idet;
  inv[4] = i4 * idet; inv[5] = i5 * idet; inv[6] = i6 * idet; inv[7] = i7 * idet;
  inv[8] = i8 * idet; inv[9] = i9 * idet; inv[10] = i10 * idet; inv[11] = i11 * idet;
}

// (K|0) * T for project2d (pose_estimator.cpp:251-268): KT = K(3x3) * T(3x4 top rows).
__device__ __forceinline__ void kt_product(const double K[9], const double T[12], double KT[12]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) KT[4 * i + j] = K[3 * i] * T[j] + K[3 * i + 1] * T[4 + j] + K[3 * i + 2] * T[8 + j];
}
__device__ __forceinline__ void kt_project(const double KT[12], double x, double y, double z, double& u, double& v) {
  double t0 = KT[0] * x + KT[1] * y + KT[2] * z + KT[3] * 1.0;
  double t1 = KT[4] * x + KT[5] * y + KT[6] * z + KT[7] * 1.0;
  double t2 = KT[8] * x + KT[9] * y + KT[10] * z + KT[11] * 1.0;
  u = t0 / t2;
  v = t1 / t2;
}

}  // namespace mpe
