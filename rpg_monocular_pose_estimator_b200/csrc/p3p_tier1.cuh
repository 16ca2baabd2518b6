// K2 tier 1 — a cheap, CONSERVATIVE pre-test that runs in front of the exact P3P solve of PoseEstimator::initialise
// (/root/reference/monocular_pose_estimator_lib/src/pose_estimator.cpp:565-702, p3p.cpp:65-286).
//
// About 99 % of the pose hypotheses of the brute-force sweep never vote (no unused detection lies within
// back_projection_pixel_tolerance_ of the back-projection of an unused LED, :669-676), yet every P3P problem pays for the
// reference's complex-arithmetic Ferrari quartic (two complex pow, six complex sqrt written after libstdc++/glibc so that
// the roots are bit-identical) and four back-substitutions.  Tier 1 answers, per problem, "certainly no vote" or "maybe":
//
//   1. the quartic's coefficients (p3p.cpp:171-185, any association, FMA allowed) are normalised and depressed;
//   2. its four roots come from a REAL factorisation into two quadratics  (t^2 + s t + u)(t^2 - s t + v)  through the
//      largest root z = s^2 of the resolvent cubic  z^3 + 2a z^2 + (a^2 - 4c) z - b^2: float seed from the closed form,
//      Newton steps in double.  A complex pair contributes its real part twice, which is what the reference uses
//      (`.real()`, p3p.cpp:276-283);
//   3. each root rho = cos(theta) with rho^2 <= 1 is back-substituted WITHOUT forming the pose: with
//      cot(alpha) = num/den the unused LED X (held in the LED triple's frame N, a table) has the camera-frame direction
//      x_c = T^T v,  v = R' X_N + (d_12 k, 0, 0)  (R' C' collapses to (-d_12 k, 0, 0)), and  K x_c = (K T^T) v  with K T^T a
//      per-detection-triple table;
//   4. the division-free test  |a_u - u a_z|^2 + |a_v - v a_z|^2 <= r^2 a_z^2  with r = tolerance + margin against every
//      unused detection.
// Tier 1 computes ACCURATE roots; the reference's Ferrari evaluation does not always (its R = -Q/2 + sqrt(Q^2/4 + P^3/27) cancels
// for Q > 0, and its 2 beta / w is 0/0-like for a nearly biquadratic quartic), and what votes is what the reference computes.
// So "maybe" is also the answer whenever the two can drift apart or the approximation cannot be trusted: more than five
// digits cancelled in R (kT1Cancel), w^2 tiny (kT1SmallW), a pair of roots closer than 1e-3 (real pair or complex pair is then
// a matter of rounding), ill-conditioned triples (same codes as the exact filter), 1 - rho^2 below 1e-4, cot(alpha) = 0/0,
// a point close to the camera plane, a factorisation whose residual is not tiny, or any non-finite intermediate.  With these
// flags the largest |rho_tier1 - rho_reference| over 1e7 unflagged hypotheses is 7.5e-9 (1e-5 px), and every deviation above
// 1e-9 found without them is explained by one of the three indicators (tests/test_cpu_k2_tier1.py prints the statistics).  Problems answered "maybe" run the unchanged exact path, so the
// histogram can only differ if tier 1 rejects a problem that the reference would have let vote; the margin (0.25 px by
// default) is ~1e6 x the deviation between tier-1 and exact back-projections observed on non-flagged problems (measured by
// tests/test_cpu_k2_tier1.py on the host build of this header and by the on/off histogram tests on the GPU).  This is a
// validated margin, not a proof: the forward error of the reference's own Ferrari evaluation is covered by the flags above
// and by measurement.  mpe_set_k2_filter(ctx, 0) switches every filter off (all-exact arm).
//
// Host + device: the header also compiles with g++ (tests build it into oracle/libtier1_check.so).
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
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MPE_HD __host__ __device__ __forceinline__
#else
#define MPE_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define T1_FMA(a, b, c) fma((a), (b), (c))
#define T1_FMAF(a, b, c) fmaf((a), (b), (c))
#else
#define T1_FMA(a, b, c) ((a) * (b) + (c))
#define T1_FMAF(a, b, c) ((a) * (b) + (c))
#endif

namespace mpe {

constexpr double kT1RootMargin2 = 1e-9;   // rho^2 > 1 + this: sqrt(1 - rho^2) is NaN in the reference as well -> no vote
constexpr double kT1NearOne = 1e-4;       // 1 - rho^2 below this: sin(theta) is ill-conditioned -> maybe
constexpr double kT1PairSep = 1e-3;       // two roots closer than this (in cos theta): real pair or complex pair is a matter of rounding -> maybe
#ifndef MPE_T1_CANCEL
#define MPE_T1_CANCEL 1e5
#endif
constexpr double kT1Cancel = MPE_T1_CANCEL;         // Q/2 over |R| in the reference's Ferrari step: more cancellation than this -> maybe
constexpr double kT1SmallW = 1e-5;        // w^2 = z below this fraction of |alpha| + sqrt|gamma|: the reference divides 2 beta by a w that is mostly rounding -> maybe
constexpr double kT1ResRel = 1e-9;        // factorisation residual |u v - c| relative to the terms -> maybe

// Tier 1 needs ~1e-12 relative accuracy, not correct rounding: reciprocal and reciprocal square root from a float seed and
// Newton steps in double — five to eight branch-free instructions instead of the ~25 (with slow-path branches) of an IEEE
// double division / square root.  Arguments outside the float range give inf / NaN, which every caller turns into "maybe".
MPE_HD double t1_rcp(double x) {
#if defined(__CUDA_ARCH__)
  double r = (double)__frcp_rn((float)x);
#else
  double r = (double)(1.0f / (float)x);
#endif
  r = T1_FMA(T1_FMA(-x, r, 1.0), r, r);          // r (2 - x r): 1e-7 -> 1e-14
  r = T1_FMA(T1_FMA(-x, r, 1.0), r, r);          // -> rounding
  return r;
}
MPE_HD double t1_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double y = (double)rsqrtf((float)x);
#else
  double y = (double)(1.0f / sqrtf((float)x));
#endif
  const double hx = 0.5 * x;
  y = T1_FMA(T1_FMA(-hx * y, y, 0.5), y, y);      // y (1.5 - x y^2 / 2)
  y = T1_FMA(T1_FMA(-hx * y, y, 0.5), y, y);
  return y;
}
MPE_HD double t1_sqrt(double x) { return (x > 0.0) ? x * t1_rsqrt(x) : ((x == 0.0) ? 0.0 : x * t1_rsqrt(x)); }

#ifdef MPE_T1_DEBUG
struct T1Debug { double al, be, ga, z, kappa, d1, d2; };
static T1Debug g_t1_debug;
#endif

struct T1Roots {
  double rho[4];
  int n;          // number of values in rho (real roots, and the real part of a complex pair twice)
  int maybe;      // 1: do not trust (caller must answer "maybe")
};

// Roots (cos theta) of the P3P quartic for (f_1, f_2, b | p_1, p_2, d_12), p3p.cpp:142-185.
MPE_HD void t1_quartic_roots(double f_1, double f_2, double b, double p_1, double p_2, double d_12, T1Roots& R) {
  R.n = 0; R.maybe = 0;
  const double F11 = f_1 * f_1, F22 = f_2 * f_2, F12 = f_1 * f_2;
  const double a = p_1, c = p_2, d = d_12;
  const double a2 = a * a, c2 = c * c, d2 = d * d, b2 = b * b, ad = a * d;
  // factors[k] / c^2 for k = 0, 1, 2, factors[3] / c and factors[4] as they are (c = p_2 != 0 for a usable triple): the same
  // polynomials as p3p.cpp:171-185, normalised by the leading one with ONE reciprocal
  const double g0 = F22 + F11 + 1.0;
  const double B = 2.0 * c * d * (T1_FMA(b, 1.0 + F22, -F12));
  const double C = T1_FMA(-F22, a2 + d2 * b2 + d2 - c2 - 2.0 * ad, T1_FMA(c2 - a2, F11, T1_FMA(2.0 * ad, 1.0 + F12 * b, -(d2 * b2 + 2.0 * a2))));
  const double D3 = 2.0 * d * (T1_FMA(a2 - ad, b, c2 * (F12 - F22 * b)));
  const double E4 = T1_FMA(F22 * c2, d2 + a2 + d2 * b2 - 2.0 * ad, T1_FMA(-2.0 * F12 * c2, ad * b, T1_FMA(c2 * F11, a2, a2 * (2.0 * ad - d2 - a2))));
  const double ic = t1_rcp(c);
  const double iA = -t1_rcp(g0) * ic * ic;            // 1 / (factors[0] / c^2)
  const double Bn = B * iA, Cn = C * iA, Dn = D3 * ic * iA, En = E4 * (ic * ic) * iA;
  // depressed quartic t^4 + al t^2 + be t + ga, x = t + sh
  const double Bn2 = Bn * Bn;
  const double al = T1_FMA(-0.375, Bn2, Cn);
  const double be = T1_FMA(Bn, T1_FMA(0.125, Bn2, -0.5 * Cn), Dn);
  const double ga = T1_FMA(Bn2, T1_FMA(-3.0 / 256.0, Bn2, Cn * (1.0 / 16.0)), T1_FMA(-0.25 * Bn, Dn, En));
  const double sh = -0.25 * Bn;
  // resolvent cubic g(z) = z^3 + 2 al z^2 + (al^2 - 4 ga) z - be^2, largest real root (>= 0 because g(0) <= 0)
  const double c2z = 2.0 * al, c1z = T1_FMA(al, al, -4.0 * ga), c0z = -(be * be);
  const double pz = T1_FMA(-1.0 / 3.0, c2z * c2z, c1z);
  const double qz = T1_FMA(c2z, T1_FMA(2.0 / 27.0, c2z * c2z, -(1.0 / 3.0) * c1z), c0z);
  const double disc = T1_FMA(0.25 * qz, qz, (1.0 / 27.0) * pz * pz * pz);
  // The reference's Ferrari evaluation forms R = -Q/2 + sqrt(Q^2/4 + P^3/27) (p3p.cpp:262; here qz = 8Q, pz = 4P, disc = 64 times
  // the radicand): for Q > 0 the two terms cancel and the roots it RETURNS drift away from the true roots of the quartic by
  // the lost digits.  Tier 1 computes accurate roots, so it stands back when more than kT1CancelDigits are lost.
#ifdef MPE_T1_DEBUG
  g_t1_debug.al = al; g_t1_debug.be = be; g_t1_debug.ga = ga; g_t1_debug.kappa = 0; g_t1_debug.z = 0; g_t1_debug.d1 = g_t1_debug.d2 = 0;
#endif
  if (qz > 0.0 && disc > 0.0) {
    const double sq = t1_sqrt(disc);
    const double Rq = ((1.0 / 27.0) * pz * pz * pz) * t1_rcp(sq + 0.5 * qz);  // = sqrt(disc) - qz/2 without the cancellation
#ifdef MPE_T1_DEBUG
    g_t1_debug.kappa = 0.5 * qz / fabs(Rq);
#endif
    if (!(fabs(Rq) * kT1Cancel >= 0.5 * qz)) { R.maybe = 1; return; }
  }
  float wseed;
  if (disc >= 0.0) {
    const float sq = sqrtf((float)disc);
    const float hq = (float)(-0.5 * qz);
    const float big = (hq >= 0.f) ? cbrtf(hq + sq) : cbrtf(hq - sq);
    wseed = big + ((big != 0.f) ? (float)(pz * (-1.0 / 3.0)) / big : 0.f);
  } else {
    const float m = 2.f * sqrtf((float)(pz * (-1.0 / 3.0)));
    float arg = 3.f * (float)qz / ((float)pz * m);
    arg = fminf(1.f, fmaxf(-1.f, arg));
    wseed = m * cosf(acosf(arg) * (1.f / 3.f));
  }
  double z = (double)wseed - c2z * (1.0 / 3.0);
  double corr = 0.0;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const double g = T1_FMA(T1_FMA(z + c2z, z, c1z), z, c0z);
    const double gp = T1_FMA(T1_FMA(3.0, z, 2.0 * c2z), z, c1z);
    corr = g * t1_rcp(gp);
    z -= corr;
  }
  if (!(z > kT1SmallW * (fabs(al) + t1_sqrt(fabs(ga)))) || !(fabs(corr) <= 1e-9 * z)) { R.maybe = 1; return; }   // also catches NaN, and be == 0 (biquadratic: z may be 0)
  const double is = t1_rsqrt(z);
  const double s = z * is;
  const double bs = be * is;
  const double u = 0.5 * (al + z - bs), v = 0.5 * (al + z + bs);
  if (!(fabs(T1_FMA(u, v, -ga)) <= kT1ResRel * (fabs(u * v) + fabs(ga) + z * z))) { R.maybe = 1; return; }
  const double d1 = T1_FMA(-4.0, u, z), d2q = T1_FMA(-4.0, v, z);
#ifdef MPE_T1_DEBUG
  g_t1_debug.z = z; g_t1_debug.d1 = d1; g_t1_debug.d2 = d2q;
#endif
  if (!(fabs(d1) > kT1PairSep * kT1PairSep) || !(fabs(d2q) > kT1PairSep * kT1PairSep)) { R.maybe = 1; return; }   // separation = sqrt(|disc|)
  if (d1 >= 0.0) { const double r = t1_sqrt(d1); R.rho[0] = 0.5 * (-s + r) + sh; R.rho[1] = 0.5 * (-s - r) + sh; }
  else { R.rho[0] = R.rho[1] = -0.5 * s + sh; }
  if (d2q >= 0.0) { const double r = t1_sqrt(d2q); R.rho[2] = 0.5 * (s + r) + sh; R.rho[3] = 0.5 * (s - r) + sh; }
  else { R.rho[2] = R.rho[3] = 0.5 * s + sh; }
  R.n = 4;
}

// Back-substitution (p3p.cpp:196-220) without divisions by f_2: cot(alpha) = num/den with both multiplied by f_2 (the sign of the
// pair is irrelevant), num = cn0 - rho cn1, den = cd0 - rho cd1.  The four constants are per problem.
struct T1Problem {
  double cn0, cn1, cd0, cd1;
  double h2_min;        // |(num, den)|^2 below this: cot(alpha) is 0/0-like -> maybe
  double b, d_12;
};
MPE_HD void t1_problem(double f_1, double f_2, double b, double p_1, double p_2, double d_12, T1Problem& Q) {
  Q.cn1 = p_2 * f_2;
  Q.cn0 = T1_FMA(d_12 * b, f_2, -(f_1 * p_1));
  Q.cd1 = f_1 * p_2;
  Q.cd0 = (p_1 - d_12) * f_2;
  const double scale = fabs(f_2) * (fabs(p_1) + fabs(p_2) + d_12 * (1.0 + fabs(b))) + fabs(f_1) * (fabs(p_1) + fabs(p_2));
  Q.h2_min = 1e-16 * scale * scale;
  Q.b = b; Q.d_12 = d_12;
}
// Root-dependent scalars
struct T1Pose {
  double rho, st;       // cos(theta), sin(theta)
  double sa, ca;        // sin(alpha), cos(alpha)
  double dk;            // d_12 * (sin(alpha) b + cos(alpha))
};
// returns 0: the reference's hypothesis is NaN (no vote), 1: usable, 2: maybe (ill-conditioned)
MPE_HD int t1_pose(double rho, const T1Problem& Q, T1Pose& P) {
  const double om = T1_FMA(-rho, rho, 1.0);
  if (om < -kT1RootMargin2) return 0;
  if (!(om >= kT1NearOne)) return 2;
  const double num = T1_FMA(-rho, Q.cn1, Q.cn0);
  const double den = T1_FMA(-rho, Q.cd1, Q.cd0);
  const double h2 = T1_FMA(num, num, den * den);
  if (!(h2 > Q.h2_min)) return 2;
  const double ih = t1_rsqrt(h2);
  P.rho = rho; P.st = om * t1_rsqrt(om);
  P.sa = fabs(den) * ih;
  P.ca = ((den >= 0.0) ? num : -num) * ih;
  P.dk = Q.d_12 * T1_FMA(P.sa, Q.b, P.ca);
  return 1;
}
MPE_HD int t1_pose(double rho, double f_1, double f_2, double b, double p_1, double p_2, double d_12, T1Pose& P) {
  T1Problem Q;
  t1_problem(f_1, f_2, b, p_1, p_2, d_12, Q);
  return t1_pose(rho, Q, P);
}

// K x_c (homogeneous pixel coordinates, not normalised) of an LED given in the triple's frame N; Mc = K [e1 e2 e3] (columns).
MPE_HD void t1_project(const T1Pose& P, const double Mc[9], double X0, double X1, double X2, double& au, double& av, double& az,
                       double& l1) {
  const double g = T1_FMA(P.rho, X1, P.st * X2);
  const double v0 = T1_FMA(-P.ca, X0, T1_FMA(-P.sa, g, P.dk));
  const double v1 = T1_FMA(P.sa, X0, -(P.ca * g));
  const double v2 = T1_FMA(-P.st, X1, P.rho * X2);
  au = T1_FMA(Mc[0], v0, T1_FMA(Mc[1], v1, Mc[2] * v2));
  av = T1_FMA(Mc[3], v0, T1_FMA(Mc[4], v1, Mc[5] * v2));
  az = T1_FMA(Mc[6], v0, T1_FMA(Mc[7], v1, Mc[8] * v2));
  l1 = fabs(v0) + fabs(v1) + fabs(v2);          // |x_c|_1 up to the rotation T (within sqrt(3))
}


// The same projection in FP32 (compile-time option MPE_T1_FP32=1; the default build keeps double): roots and the root-dependent scalars stay in double, the unused LED's
// homogeneous pixel coordinates and the comparisons against the unused detections are single precision.  Error budget against the
// 0.25 px margin: |v| ~ 1 m and |Mc| <= ~1e3 give |a| ~ 1e3 with a rounding error of a few 1e-4, i.e. < 1e-3 px after the
// (implicit) division by a_z >= 1e-3 |v|_1 ... in practice a_z ~ |v|; the products u a_z add 752 * 6e-8.  Measured by
// tests/test_cpu_k2_tier1.py (host build, same float arithmetic): the deviation from the exact back-projection stays below
// 3e-3 px on unflagged problems.  Halves the registers tier 1 holds (K T^T, the LED coordinates, the detections) and moves
// 60 % of its arithmetic to the otherwise idle FP32 pipe.
struct T1PoseF { float rho, st, sa, ca, dk; };
MPE_HD T1PoseF t1_pose_f(const T1Pose& P) {
  T1PoseF F; F.rho = (float)P.rho; F.st = (float)P.st; F.sa = (float)P.sa; F.ca = (float)P.ca; F.dk = (float)P.dk; return F;
}
MPE_HD void t1_project_f(const T1PoseF& P, const float Mc[9], float X0, float X1, float X2, float& au, float& av, float& az, float& l1) {
  const float g = T1_FMAF(P.rho, X1, P.st * X2);
  const float v0 = T1_FMAF(-P.ca, X0, T1_FMAF(-P.sa, g, P.dk));
  const float v1 = T1_FMAF(P.sa, X0, -(P.ca * g));
  const float v2 = T1_FMAF(-P.st, X1, P.rho * X2);
  au = T1_FMAF(Mc[0], v0, T1_FMAF(Mc[1], v1, Mc[2] * v2));
  av = T1_FMAF(Mc[3], v0, T1_FMAF(Mc[4], v1, Mc[5] * v2));
  az = T1_FMAF(Mc[6], v0, T1_FMAF(Mc[7], v1, Mc[8] * v2));
  l1 = fabsf(v0) + fabsf(v1) + fabsf(v2);
}

}  // namespace mpe
