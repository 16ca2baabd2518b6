// TEST INFRASTRUCTURE — CPU oracle for the pose side of the hot path.
//
// This file is a from-scratch C++17 restatement (no Eigen, no OpenCV) of the reference's
// pose arithmetic.  It is NOT part of the product: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product path is the
// CUDA library behind include/mpe_b200.h and never calls into this file.
//
// PARITY STATUS: "parity unpinned" by the reference — the reference ships no tests, golden
// vectors or fixtures and cannot be compiled here (no Eigen/OpenCV/ROS).
// What pins this file instead:
//   * solveQuartic/computePoses use std::complex<double> exactly as the reference does, so the
//     libstdc++/glibc/libgcc complex pow/sqrt/div behaviour is inherited, not re-implemented;
//   * oracle/_ref/libref_p3p.so (built by oracle/Makefile from the UNMODIFIED
//     /root/reference/.../src/p3p.cpp against a tiny Eigen stand-in) is compared bit-for-bit
//     with p3p_compute_poses() below in tests/test_oracle_p3p_ref.py when it is available;
//   * tests/golden/*.npz hold outputs of this oracle for seeded scenes (generator committed).
//
// Build flags mirror the reference (monocular_pose_estimator_lib/CMakeLists.txt:5-6):
//   g++ -std=c++17 -O3 -ffp-contract=off   (no -march=native => no FMA contraction)
//
// Every function cites the reference lines it follows.  L/ = monocular_pose_estimator_lib/.
// Matrices are row-major double[ ] here; the C API converts where stated.

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
#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

using V2 = std::array<double, 2>;
using V3 = std::array<double, 3>;
using V4 = std::array<double, 4>;
struct M3 { double m[3][3]; };
struct M4 { double m[4][4]; };
struct M34 { double m[3][4]; };

const double kInf = std::numeric_limits<double>::infinity();

// ---------------------------------------------------------------- small linear algebra
inline V3 sub(const V3& a, const V3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline V3 add(const V3& a, const V3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
inline V3 cross(const V3& a, const V3& b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline double dot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }
inline V3 divs(const V3& a, double s) { return {a[0] / s, a[1] / s, a[2] / s}; }
inline V3 mul(const M3& A, const V3& v) {
  V3 r;
  for (int i = 0; i < 3; ++i) r[i] = A.m[i][0] * v[0] + A.m[i][1] * v[1] + A.m[i][2] * v[2];
  return r;
}
inline M3 mul(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
  return C;
}
inline M3 transpose(const M3& A) {
  M3 T;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T.m[i][j] = A.m[j][i];
  return T;
}
inline M3 rows(const V3& a, const V3& b, const V3& c) {
  M3 T;
  for (int j = 0; j < 3; ++j) { T.m[0][j] = a[j]; T.m[1][j] = b[j]; T.m[2][j] = c[j]; }
  return T;
}
inline M4 identity4() {
  M4 I;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) I.m[i][j] = (i == j) ? 1.0 : 0.0;
  return I;
}
inline M4 mul(const M4& A, const M4& B) {
  M4 C;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = A.m[i][0] * B.m[0][j];
      for (int k = 1; k < 4; ++k) s += A.m[i][k] * B.m[k][j];
      C.m[i][j] = s;
    }
  return C;
}
inline V4 mul(const M4& A, const V4& v) {
  V4 r;
  for (int i = 0; i < 4; ++i) {
    double s = A.m[i][0] * v[0];
    for (int k = 1; k < 4; ++k) s += A.m[i][k] * v[k];
    r[i] = s;
  }
  return r;
}

// General 4x4 inverse by cofactors (Eigen's fixed-size 4x4 inverse() is also cofactor based;
// used at L/src/pose_estimator.cpp:486,516,660 and :235).  No rigid-body shortcut is taken.
M4 inverse4(const M4& A) {
  const double* a = &A.m[0][0];
  double inv[16];
  inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
  inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
  inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
  inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
  inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
  inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
  inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
  inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
  inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
  inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
  inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
  inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
  inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
  inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
  inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
  inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
  double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
  double idet = 1.0 / det;
  M4 R;
  for (int i = 0; i < 16; ++i) (&R.m[0][0])[i] = inv[i] * idet;
  return R;
}

// ---------------------------------------------------------------- P3P (L/src/p3p.cpp)

// L/src/p3p.cpp:238-286.  std::complex<double> arithmetic is used exactly as in the
// reference so that libstdc++'s pow(complex,double) / sqrt(complex) / operator/ semantics
// (polar form for non-positive-real bases, __divdc3 division) are inherited.
void solve_quartic(const double factors[5], double real_roots[4]) {
  double A = factors[0], B = factors[1], C = factors[2], D = factors[3], E = factors[4];
  double A_pw2 = A * A, B_pw2 = B * B;
  double A_pw3 = A_pw2 * A, B_pw3 = B_pw2 * B;
  double A_pw4 = A_pw3 * A, B_pw4 = B_pw3 * B;

  double alpha = -3 * B_pw2 / (8 * A_pw2) + C / A;
  double beta = B_pw3 / (8 * A_pw3) - B * C / (2 * A_pw2) + D / A;
  double gamma = -3 * B_pw4 / (256 * A_pw4) + B_pw2 * C / (16 * A_pw3) - B * D / (4 * A_pw2) + E / A;

  double alpha_pw2 = alpha * alpha;
  double alpha_pw3 = alpha_pw2 * alpha;

  std::complex<double> P(-alpha_pw2 / 12 - gamma, 0);
  std::complex<double> Q(-alpha_pw3 / 108 + alpha * gamma / 3 - std::pow(beta, 2) / 8, 0);
  std::complex<double> R = -Q / 2.0 + std::sqrt(std::pow(Q, 2.0) / 4.0 + std::pow(P, 3.0) / 27.0);

  std::complex<double> U = std::pow(R, (1.0 / 3.0));
  std::complex<double> y;
  if (U.real() == 0)
    y = -5.0 * alpha / 6.0 - std::pow(Q, (1.0 / 3.0));
  else
    y = -5.0 * alpha / 6.0 - P / (3.0 * U) + U;

  std::complex<double> w = std::sqrt(alpha + 2.0 * y);
  std::complex<double> temp;
  temp = -B / (4.0 * A) + 0.5 * (w + std::sqrt(-(3.0 * alpha + 2.0 * y + 2.0 * beta / w)));
  real_roots[0] = temp.real();
  temp = -B / (4.0 * A) + 0.5 * (w - std::sqrt(-(3.0 * alpha + 2.0 * y + 2.0 * beta / w)));
  real_roots[1] = temp.real();
  temp = -B / (4.0 * A) + 0.5 * (-w + std::sqrt(-(3.0 * alpha + 2.0 * y - 2.0 * beta / w)));
  real_roots[2] = temp.real();
  temp = -B / (4.0 * A) + 0.5 * (-w - std::sqrt(-(3.0 * alpha + 2.0 * y - 2.0 * beta / w)));
  real_roots[3] = temp.real();
}

// L/src/p3p.cpp:65-236.  f[k], Pw[k] are the k-th columns of feature_vectors / world_points.
// Returns 0, or -1 when the world points are colinear (:77-80).  sol[i] is [R | C] (3x4).
int p3p_compute_poses(const V3 f_in[3], const V3 P_in[3], M34 sol[4]) {
  V3 P1 = P_in[0], P2 = P_in[1], P3 = P_in[2];
  V3 temp1 = sub(P2, P1), temp2 = sub(P3, P1);
  if (norm(cross(temp1, temp2)) == 0) return -1;  // :77

  V3 f1 = f_in[0], f2 = f_in[1], f3 = f_in[2];
  V3 e1 = f1;
  V3 e3 = cross(f1, f2);
  e3 = divs(e3, norm(e3));
  V3 e2 = cross(e3, e1);
  M3 T = rows(e1, e2, e3);
  f3 = mul(T, f3);

  if (f3[2] > 0) {  // :101-121
    f1 = f_in[1]; f2 = f_in[0]; f3 = f_in[2];
    e1 = f1;
    e3 = cross(f1, f2);
    e3 = divs(e3, norm(e3));
    e2 = cross(e3, e1);
    T = rows(e1, e2, e3);
    f3 = mul(T, f3);
    P1 = P_in[1]; P2 = P_in[0]; P3 = P_in[2];
  }

  V3 n1 = sub(P2, P1);  // :124-136
  n1 = divs(n1, norm(n1));
  V3 n3 = cross(n1, sub(P3, P1));
  n3 = divs(n3, norm(n3));
  V3 n2 = cross(n3, n1);
  M3 N = rows(n1, n2, n3);

  P3 = mul(N, sub(P3, P1));  // :139
  double d_12 = norm(sub(P2, P1));
  double f_1 = f3[0] / f3[2];
  double f_2 = f3[1] / f3[2];
  double p_1 = P3[0];
  double p_2 = P3[1];

  double cos_beta = dot(f1, f2);
  double b = 1 / (1 - std::pow(cos_beta, 2)) - 1;
  if (cos_beta < 0) b = -std::sqrt(b); else b = std::sqrt(b);

  double f_1_pw2 = std::pow(f_1, 2), f_2_pw2 = std::pow(f_2, 2);
  double p_1_pw2 = std::pow(p_1, 2), p_1_pw3 = p_1_pw2 * p_1, p_1_pw4 = p_1_pw3 * p_1;
  double p_2_pw2 = std::pow(p_2, 2), p_2_pw3 = p_2_pw2 * p_2, p_2_pw4 = p_2_pw3 * p_2;
  double d_12_pw2 = std::pow(d_12, 2), b_pw2 = std::pow(b, 2);

  double factors[5];  // :171-185, expression order kept
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

  double realRoots[4];
  solve_quartic(factors, realRoots);

  M3 Nt = transpose(N);
  for (int i = 0; i < 4; ++i) {  // :193-233
    double cot_alpha = (-f_1 * p_1 / f_2 - realRoots[i] * p_2 + d_12 * b) / (-f_1 * realRoots[i] * p_2 / f_2 + p_1 - d_12);
    double cos_theta = realRoots[i];
    double sin_theta = std::sqrt(1 - std::pow(realRoots[i], 2));
    double sin_alpha = std::sqrt(1 / (std::pow(cot_alpha, 2) + 1));
    double cos_alpha = std::sqrt(1 - std::pow(sin_alpha, 2));
    if (cot_alpha < 0) cos_alpha = -cos_alpha;

    V3 C;
    C[0] = d_12 * cos_alpha * (sin_alpha * b + cos_alpha);
    C[1] = cos_theta * d_12 * sin_alpha * (sin_alpha * b + cos_alpha);
    C[2] = sin_theta * d_12 * sin_alpha * (sin_alpha * b + cos_alpha);
    C = add(P1, mul(Nt, C));

    M3 R;
    R.m[0][0] = -cos_alpha;            R.m[0][1] = -sin_alpha * cos_theta; R.m[0][2] = -sin_alpha * sin_theta;
    R.m[1][0] = sin_alpha;             R.m[1][1] = -cos_alpha * cos_theta; R.m[1][2] = -cos_alpha * sin_theta;
    R.m[2][0] = 0;                     R.m[2][1] = -sin_theta;             R.m[2][2] = cos_theta;
    R = mul(mul(Nt, transpose(R)), T);  // (N^T * R^T) * T, left to right as Eigen evaluates

    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) sol[i].m[r][c] = R.m[r][c];
      sol[i].m[r][3] = C[r];
    }
  }
  return 0;
}

// ---------------------------------------------------------------- index tables (L/src/combinations.cpp)

// Row order of Combinations::combinationsNoReplacement(N,3) (L/src/combinations.cpp:60-125):
// lexicographic; values 1-based there, 0-based here.
std::vector<std::array<unsigned, 3>> combinations3(unsigned N) {
  std::vector<std::array<unsigned, 3>> out;
  for (unsigned a = 0; a < N; ++a)
    for (unsigned b = a + 1; b < N; ++b)
      for (unsigned c = b + 1; c < N; ++c) out.push_back({a, b, c});
  return out;
}

// Row order of Combinations::permutationsNoReplacement(N,3) (L/src/combinations.cpp:127-208 with
// permutations(3) from :210-244): for each lexicographic combination (a<b<c) the six rows
// [c b a],[c a b],[b c a],[b a c],[a b c],[a c b].  0-based here.
std::vector<std::array<unsigned, 3>> permutations3(unsigned N) {
  std::vector<std::array<unsigned, 3>> out;
  for (const auto& k : combinations3(N)) {
    unsigned a = k[0], b = k[1], c = k[2];
    out.push_back({c, b, a}); out.push_back({c, a, b}); out.push_back({b, c, a});
    out.push_back({b, a, c}); out.push_back({a, b, c}); out.push_back({a, c, b});
  }
  return out;
}

unsigned factorial_u(int N) { return (N == 1 || N == 0) ? 1u : factorial_u(N - 1) * (unsigned)N; }  // combinations.cpp:34-40 (unsigned wrap kept)
unsigned num_combinations(unsigned N, unsigned K) { return factorial_u((int)N) / (factorial_u((int)K) * factorial_u((int)(N - K))); }  // :42-45

// ---------------------------------------------------------------- 6x6 and 3x3 dense helpers

// Solve A x = b for symmetric A: LDL^T with diagonal pivoting as Eigen's LDLT documents it (L/src/pose_estimator.cpp:778),
// operation for operation what oracle/eigen_shim's LDLT does, so that the Gauss-Newton iteration count (exit test at the
// rounding floor, :786) equals the count of the reference sources built against that stand-in:
//   * lower triangle only, left-looking: column k is formed from the original entries and the finished columns 0..k-1;
//   * the pivot of step k is the largest |diagonal| among the NOT YET UPDATED entries k..5 (first one on ties);
//   * solve: P, unit-lower forward substitution, D (a pivot below DBL_MIN gives 0), backward substitution, P^T.
void ldlt_solve6(const double Ain[36], const double bin[6], double x[6]) {
  double m[6][6];
  int tr[6];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) m[i][j] = Ain[i * 6 + j];
  for (int k = 0; k < 6; ++k) {
    int big = k; double best = std::fabs(m[k][k]);
    for (int i = k + 1; i < 6; ++i) if (std::fabs(m[i][i]) > best) { best = std::fabs(m[i][i]); big = i; }
    tr[k] = big;
    if (big != k) {
      for (int j = 0; j < k; ++j) std::swap(m[k][j], m[big][j]);
      for (int i = big + 1; i < 6; ++i) std::swap(m[i][k], m[i][big]);
      std::swap(m[k][k], m[big][big]);
      for (int i = k + 1; i < big; ++i) std::swap(m[i][k], m[big][i]);
    }
    if (k > 0) {
      double temp[6];
      for (int j = 0; j < k; ++j) temp[j] = m[j][j] * m[k][j];
      double s = m[k][0] * temp[0];
      for (int j = 1; j < k; ++j) s += m[k][j] * temp[j];
      m[k][k] -= s;
      for (int i = k + 1; i < 6; ++i) {
        double t = m[i][0] * temp[0];
        for (int j = 1; j < k; ++j) t += m[i][j] * temp[j];
        m[i][k] -= t;
      }
    }
    const double akk = m[k][k];
    if (std::fabs(akk) > 0.0) for (int i = k + 1; i < 6; ++i) m[i][k] /= akk;
  }
  for (int i = 0; i < 6; ++i) x[i] = bin[i];
  for (int k = 0; k < 6; ++k) if (tr[k] != k) std::swap(x[k], x[tr[k]]);
  for (int j = 0; j < 6; ++j) for (int i = j + 1; i < 6; ++i) x[i] -= m[i][j] * x[j];
  for (int i = 0; i < 6; ++i) { if (std::fabs(m[i][i]) > std::numeric_limits<double>::min()) x[i] /= m[i][i]; else x[i] = 0.0; }
  for (int j = 5; j >= 0; --j) for (int i = 0; i < j; ++i) x[i] -= m[j][i] * x[j];
  for (int k = 5; k >= 0; --k) if (tr[k] != k) std::swap(x[k], x[tr[k]]);
}

// General 6x6 inverse, LU with partial pivoting (Eigen's inverse() for size > 4 goes through
// PartialPivLU; L/src/pose_estimator.cpp:790).
void inverse6(const double Ain[36], double out[36]) {
  double a[6][12];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { a[i][j] = Ain[i * 6 + j]; a[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  for (int k = 0; k < 6; ++k) {
    int p = k; double best = std::fabs(a[k][k]);
    for (int i = k + 1; i < 6; ++i) if (std::fabs(a[i][k]) > best) { best = std::fabs(a[i][k]); p = i; }
    if (p != k) for (int j = 0; j < 12; ++j) std::swap(a[k][j], a[p][j]);
    double piv = a[k][k];
    for (int j = 0; j < 12; ++j) a[k][j] /= piv;
    for (int i = 0; i < 6; ++i) if (i != k) {
      double f = a[i][k];
      if (f != 0) for (int j = 0; j < 12; ++j) a[i][j] -= f * a[k][j];
    }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) out[i * 6 + j] = a[i][6 + j];
}

// 3x3 SVD A = U S V^T by one-sided Jacobi on columns (any SVD will do for Kabsch: the sign /
// ordering ambiguities cancel in V U^T; L/src/pose_estimator.cpp:916-922).
void svd3(const M3& Ain, M3& U, M3& V) {
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { a[i][j] = Ain.m[i][j]; v[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
      double alpha = 0, beta = 0, gamma = 0;
      for (int i = 0; i < 3; ++i) { alpha += a[i][p] * a[i][p]; beta += a[i][q] * a[i][q]; gamma += a[i][p] * a[i][q]; }
      if (gamma == 0) continue;
      off = std::max(off, std::fabs(gamma) / std::sqrt(alpha * beta));
      double zeta = (beta - alpha) / (2.0 * gamma);
      double t = ((zeta >= 0) ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
      double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
      for (int i = 0; i < 3; ++i) {
        double ap = a[i][p], aq = a[i][q];
        a[i][p] = c * ap - s * aq; a[i][q] = s * ap + c * aq;
        double vp = v[i][p], vq = v[i][q];
        v[i][p] = c * vp - s * vq; v[i][q] = s * vp + c * vq;
      }
    }
    if (off < 1e-16) break;
  }
  // columns of a are U*S; normalise.  A rank-deficient column is completed by a cross product.
  double sv[3];
  for (int j = 0; j < 3; ++j) sv[j] = std::sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int x, int y) { return sv[x] > sv[y]; });
  for (int jj = 0; jj < 3; ++jj) {
    int j = order[jj];
    for (int i = 0; i < 3; ++i) { V.m[i][jj] = v[i][j]; U.m[i][jj] = (sv[j] > 0) ? a[i][j] / sv[j] : 0.0; }
  }
  if (!(sv[order[2]] > 1e-300 * sv[order[0]])) {
    V3 u0{U.m[0][0], U.m[1][0], U.m[2][0]}, u1{U.m[0][1], U.m[1][1], U.m[2][1]};
    V3 u2 = cross(u0, u1);
    for (int i = 0; i < 3; ++i) U.m[i][2] = u2[i];
  }
}

// ---------------------------------------------------------------- the estimator

struct Oracle {
  // L/include/.../pose_estimator.h:56-91 (members)
  M4 current_pose, previous_pose, predicted_pose;
  double pose_covariance[36];
  double current_time = 0, previous_time = 0, predicted_time = 0;
  std::vector<V4> object_points;
  std::vector<V2> image_points, predicted_pixel_positions;
  std::vector<V3> image_vectors;
  std::vector<std::array<unsigned, 2>> correspondences;  // (LED, detection), 1-based, 0 = none
  double back_projection_pixel_tolerance = 3, nearest_neighbour_pixel_tolerance = 5;  // pose_estimator.cpp:36-39
  double certainty_threshold = 0.75, valid_correspondence_threshold = 0.7;
  unsigned histogram_threshold = 0;
  unsigned it_since_initialized = 0;
  double K[3][3];
  std::vector<double> D;
  std::vector<unsigned> last_hist;  // n_det x n_obj row-major copy of hist_corr before decoding
  unsigned last_gn_iterations = 0;
  // instrumentation (not in the reference): exact FP64 op-free counters for DESIGN.md
  unsigned long long n_p3p = 0, n_finite = 0, n_voting = 0;

  Oracle() {
    current_pose = previous_pose = predicted_pose = identity4();
    std::memset(pose_covariance, 0, sizeof(pose_covariance));
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) K[i][j] = (i == j) ? 1.0 : 0.0;
  }

  // pose_estimator.cpp:50-55
  void set_marker_positions(const double* xyz, unsigned n) {
    object_points.resize(n);
    for (unsigned i = 0; i < n; ++i) object_points[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 1.0};
    predicted_pixel_positions.assign(n, V2{0, 0});
    histogram_threshold = num_combinations(n, 3);
  }

  // pose_estimator.cpp:288-301
  void calculate_image_vectors() {
    image_vectors.resize(image_points.size());
    for (size_t i = 0; i < image_points.size(); ++i) {
      V3 v;
      v[0] = (image_points[i][0] - K[0][2]) / K[0][0];
      v[1] = (image_points[i][1] - K[1][2]) / K[1][1];
      v[2] = 1;
      image_vectors[i] = divs(v, norm(v));
    }
  }
  void set_image_points(const double* pts, unsigned n) {  // :166-170
    image_points.resize(n);
    for (unsigned i = 0; i < n; ++i) image_points[i] = {pts[2 * i], pts[2 * i + 1]};
    calculate_image_vectors();
  }

  // pose_estimator.cpp:251-268: (K|0) * T first, then * p, then divide by z.
  V2 project2d(const V4& point, const M4& T) const {
    double KT[3][4];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) {
        double s = K[i][0] * T.m[0][j];
        s += K[i][1] * T.m[1][j];
        s += K[i][2] * T.m[2][j];
        s += 0.0 * T.m[3][j];
        KT[i][j] = s;
      }
    double t[3];
    for (int i = 0; i < 3; ++i) {
      double s = KT[i][0] * point[0];
      for (int k = 1; k < 4; ++k) s += KT[i][k] * point[k];
      t[i] = s;
    }
    return {t[0] / t[2], t[1] / t[2]};
  }

  static bool is_finite(const M4& x) {  // :856-860: ((x-x)==(x-x)).all()
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double d = x.m[i][j] - x.m[i][j]; if (!(d == d)) return false; }
    return true;
  }
  static double square_dist(const V2& a, const V2& b) { double dx = a[0] - b[0], dy = a[1] - b[1]; return dx * dx + dy * dy; }  // :850-854

  // :862-906.  pairs[i] = 1-based index into b of the nearest point (0 if b empty), strict '<'.
  static void min_distances_and_pairs(const std::vector<V2>& a, const std::vector<V2>& b, std::vector<unsigned>& pairs, std::vector<double>& min_d) {
    pairs.assign(a.size(), 0);
    min_d.assign(a.size(), 0);
    for (size_t i = 0; i < a.size(); ++i) {
      double best = kInf;
      for (size_t j = 0; j < b.size(); ++j) {
        double d2 = square_dist(a[i], b[j]);
        if (d2 < best) { best = d2; pairs[i] = (unsigned)j + 1; }
      }
      min_d[i] = std::sqrt(best);
    }
  }

  // :303-342.  distances(i,j): i image pts (rows), j object pts (cols); minCoeff visits
  // column-major and keeps the first strict minimum.
  double squared_error_and_certainty(const std::vector<V2>& image_pts, const std::vector<V2>& object_pts, double& certainty) const {
    size_t ni = image_pts.size(), no = object_pts.size();
    std::vector<double> d(ni * no);
    for (size_t i = 0; i < ni; ++i) for (size_t j = 0; j < no; ++j) d[i * no + j] = std::sqrt(square_dist(image_pts[i], object_pts[j]));
    double squared_error = 0; unsigned num = 0;
    for (size_t it = 1; it <= std::min(ni, no); ++it) {
      double mv = d[0]; size_t ri = 0, ci = 0;
      for (size_t j = 0; j < no; ++j) for (size_t i = 0; i < ni; ++i) if (d[i * no + j] < mv) { mv = d[i * no + j]; ri = i; ci = j; }
      if (mv <= back_projection_pixel_tolerance) {
        squared_error += std::pow(d[ri * no + ci], 2);
        ++num;
        for (size_t j = 0; j < no; ++j) d[ri * no + j] = kInf;
        for (size_t i = 0; i < ni; ++i) d[i * no + ci] = kInf;
      } else break;
    }
    certainty = (double)num / (double)no;
    return squared_error;
  }

  // :344-370.  hist is n_det x n_obj (row = detection, col = LED); maxCoeff = first strict max
  // in column-major order; only the chosen column is cleared.
  void correspondences_from_histogram(std::vector<unsigned>& hist, unsigned n_det, unsigned n_obj) {
    correspondences.clear();
    for (unsigned j = 0; j < n_obj; ++j) {
      unsigned mv = hist[0], ri = 0, ci = 0;
      for (unsigned c = 0; c < n_obj; ++c) for (unsigned r = 0; r < n_det; ++r) if (hist[r * n_obj + c] > mv) { mv = hist[r * n_obj + c]; ri = r; ci = c; }
      if (mv < histogram_threshold) break;
      correspondences.push_back({ci + 1, ri + 1});
      for (unsigned r = 0; r < n_det; ++r) hist[r * n_obj + ci] = 0;
    }
  }

  // :372-392
  void find_correspondences() {
    std::vector<unsigned> pairs; std::vector<double> md;
    min_distances_and_pairs(predicted_pixel_positions, image_points, pairs, md);
    correspondences.clear();
    for (size_t i = 0; i < pairs.size(); ++i)
      if (md[i] <= nearest_neighbour_pixel_tolerance) correspondences.push_back({(unsigned)i + 1, pairs[i]});
  }

  static M4 from34(const M34& s) {
    M4 H = identity4();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) H.m[r][c] = s.m[r][c];
    return H;
  }

  // :908-930 (Kabsch, no reflection fix)
  static M4 compute_transformation(const std::vector<V3>& obj, const std::vector<V3>& rep) {
    size_t n = obj.size();
    V3 mo{0, 0, 0}, mr{0, 0, 0};
    for (size_t i = 0; i < n; ++i) { mo = add(mo, obj[i]); mr = add(mr, rep[i]); }
    mo = divs(mo, (double)n); mr = divs(mr, (double)n);
    M3 Hm;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
      double s = 0;
      for (size_t i = 0; i < n; ++i) s += (obj[i][r] - mo[r]) * (rep[i][c] - mr[c]);
      Hm.m[r][c] = s;
    }
    M3 U, V; svd3(Hm, U, V);
    M3 R = mul(V, transpose(U));
    V3 t = sub(mr, mul(R, mo));
    M4 T = identity4();
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T.m[r][c] = R.m[r][c]; T.m[r][3] = t[r]; }
    return T;
  }

  // :394-542
  unsigned check_correspondences() {
    unsigned nc = (unsigned)correspondences.size();
    if (nc < 4) return 0;
    size_t n_obj = object_points.size();
    std::vector<V4> mean_rep(n_obj, V4{0, 0, 0, 0});
    auto combos = combinations3(nc);
    unsigned N = (unsigned)combos.size();
    unsigned num_valid = 0;
    for (unsigned i = 0; i < N; ++i) {
      V3 fv[3], wp[3];
      for (int k = 0; k < 3; ++k) {
        const V4& op = object_points[correspondences[combos[i][k]][0] - 1];
        wp[k] = {op[0], op[1], op[2]};
        fv[k] = image_vectors[correspondences[combos[i][k]][1] - 1];
      }
      std::vector<V2> unused_im; std::vector<V4> unused_obj;
      for (unsigned l = 0; l < nc; ++l) {
        if (l == combos[i][0] || l == combos[i][1] || l == combos[i][2]) continue;
        unused_obj.push_back(object_points[correspondences[l][0] - 1]);
        unused_im.push_back(image_points[correspondences[l][1] - 1]);
      }
      M34 sol[4];
      if (p3p_compute_poses(fv, wp, sol) != 0) continue;
      double min_sq = kInf; unsigned best = 0; bool found = false;
      for (unsigned j = 0; j < 4; ++j) {
        M4 H = from34(sol[j]);
        if (!is_finite(H)) continue;
        M4 Hi = inverse4(H);
        std::vector<V2> back(unused_obj.size());
        for (size_t ii = 0; ii < unused_obj.size(); ++ii) back[ii] = project2d(unused_obj[ii], Hi);
        double certainty;
        double sq = squared_error_and_certainty(unused_im, back, certainty);
        if (certainty >= certainty_threshold) {
          found = true;
          if (sq < min_sq) { min_sq = sq; best = j; }
        }
      }
      if (found) {
        ++num_valid;
        M4 Hi = inverse4(from34(sol[best]));
        for (size_t jj = 0; jj < n_obj; ++jj) {
          V4 p = mul(Hi, object_points[jj]);
          for (int k = 0; k < 4; ++k) mean_rep[jj][k] = mean_rep[jj][k] + p[k];
        }
      }
    }
    if ((double)num_valid / N >= valid_correspondence_threshold) {
      std::vector<V3> obj(n_obj), rep(n_obj);
      for (size_t k = 0; k < n_obj; ++k) {
        obj[k] = {object_points[k][0], object_points[k][1], object_points[k][2]};
        rep[k] = {mean_rep[k][0] / num_valid, mean_rep[k][1] / num_valid, mean_rep[k][2] / num_valid};
      }
      predicted_pose = compute_transformation(obj, rep);
      return 1;
    }
    return 0;
  }

  // :544-721
  unsigned initialise() {
    unsigned n_det = (unsigned)image_points.size(), n_obj = (unsigned)object_points.size();
    auto seen = combinations3(n_det);
    auto perms = permutations3(n_obj);
    std::vector<unsigned> hist(n_det * n_obj, 0);
    for (const auto& sc : seen) {
      V3 fv[3] = {image_vectors[sc[0]], image_vectors[sc[1]], image_vectors[sc[2]]};
      std::vector<V2> unused_im; std::vector<unsigned> unused_im_idx;
      for (unsigned kk = 0; kk < n_det; ++kk) if (kk != sc[0] && kk != sc[1] && kk != sc[2]) { unused_im.push_back(image_points[kk]); unused_im_idx.push_back(kk); }
      for (const auto& pm : perms) {
        V3 wp[3];
        for (int k = 0; k < 3; ++k) { const V4& op = object_points[pm[k]]; wp[k] = {op[0], op[1], op[2]}; }
        M34 sol[4];
        ++n_p3p;
        if (p3p_compute_poses(fv, wp, sol) != 0) continue;
        std::vector<V4> unused_obj; std::vector<unsigned> unused_obj_idx;
        for (unsigned ll = 0; ll < n_obj; ++ll) if (ll != pm[0] && ll != pm[1] && ll != pm[2]) { unused_obj.push_back(object_points[ll]); unused_obj_idx.push_back(ll); }
        for (unsigned k = 0; k < 4; ++k) {
          M4 H = from34(sol[k]);
          if (!is_finite(H)) continue;
          ++n_finite;
          M4 Hi = inverse4(H);
          std::vector<V2> back(unused_obj.size());
          for (size_t m = 0; m < unused_obj.size(); ++m) back[m] = project2d(unused_obj[m], Hi);
          std::vector<unsigned> pairs; std::vector<double> md;
          min_distances_and_pairs(unused_im, back, pairs, md);
          unsigned cnt = 0;
          for (double d : md) if (d < back_projection_pixel_tolerance) ++cnt;
          if (cnt > 0) {
            ++n_voting;
            for (int mm = 0; mm < 3; ++mm) hist[sc[mm] * n_obj + pm[mm]] += 1;
            for (size_t nn = 0; nn < md.size(); ++nn)
              if (md[nn] < back_projection_pixel_tolerance) hist[unused_im_idx[nn] * n_obj + unused_obj_idx[pairs[nn] - 1]] += 1;
          }
        }
      }
    }
    last_hist = hist;
    bool all_zero = true;
    for (unsigned v : hist) if (v != 0) { all_zero = false; break; }
    if (all_zero) return 0;
    correspondences_from_histogram(hist, n_det, n_obj);
    return check_correspondences() == 1 ? 1 : 0;
  }

  // :1066-1071, :962-994
  static M3 skew(const V3& w) { M3 O{{{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}}}; return O; }
  static M4 exponential_map(const double twist[6]) {
    V3 upsilon{twist[0], twist[1], twist[2]}, omega{twist[3], twist[4], twist[5]};
    double theta = norm(omega), theta_squared = theta * theta;
    M3 Omega = skew(omega), Omega2 = mul(Omega, Omega), rot, V;
    if (theta == 0) {
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rot.m[i][j] = V.m[i][j] = (i == j) ? 1.0 : 0.0;
    } else {
      double s = std::sin(theta), c = std::cos(theta);
      double kv1 = (1 - c) / (theta_squared), kv2 = (theta - s) / (theta_squared * theta);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double I = (i == j) ? 1.0 : 0.0;
        rot.m[i][j] = I + Omega.m[i][j] / theta * s + Omega2.m[i][j] / theta_squared * (1 - c);
        V.m[i][j] = I + kv1 * Omega.m[i][j] + kv2 * Omega2.m[i][j];
      }
    }
    V3 t = mul(V, upsilon);
    M4 T = identity4();
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T.m[r][c] = rot.m[r][c]; T.m[r][3] = t[r]; }
    return T;
  }

  // :996-1064
  static void logarithm_map(const M4& trans, double xi[6]) {
    M3 R; V3 t;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R.m[r][c] = trans.m[r][c]; t[r] = trans.m[r][3]; }
    M3 w_hat; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) w_hat.m[i][j] = 0;
    double phi = 0;
    // R.isApprox(I, 1e-10): ||R-I||_F^2 <= 1e-20 * min(||R||_F^2, ||I||_F^2)
    double dn = 0, rn = 0;
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) { double I = (i == j) ? 1.0 : 0.0; dn += (R.m[i][j] - I) * (R.m[i][j] - I); rn += R.m[i][j] * R.m[i][j]; }  // column-major, as Eigen stores R
    bool approx_identity = dn <= 1e-10 * 1e-10 * std::min(rn, 3.0);
    if (!approx_identity) {
      double temp = (R.m[0][0] + R.m[1][1] + R.m[2][2] - 1) / 2;
      if (temp > 1) temp = 1; else if (temp < -1) temp = -1;
      phi = std::acos(temp);
      if (phi != 0) {
        double s2 = 2 * std::sin(phi);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) w_hat.m[i][j] = (R.m[i][j] - R.m[j][i]) / s2 * phi;
      }
    }
    V3 w{w_hat.m[2][1], w_hat.m[0][2], w_hat.m[1][0]};
    double w_norm = norm(w);
    M3 A_inv;
    // t.isApproxToConstant(0, 1e-10) is true only for t == 0 exactly (|t_i - 0| <= 1e-10*min(|t_i|,0))
    bool t_zero = (t[0] == 0 && t[1] == 0 && t[2] == 0);
    if (t_zero) {
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A_inv.m[i][j] = 0;
    } else if (w_norm == 0 || std::sin(w_norm) == 0) {
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A_inv.m[i][j] = (i == j) ? 1.0 : 0.0;
    } else {
      double k = (2 * std::sin(w_norm) - w_norm * (1 + std::cos(w_norm))) / (2 * w_norm * w_norm * std::sin(w_norm));
      // `I - w_hat / 2 + k * w_hat * w_hat` parses as (I - w_hat/2) + ((k * w_hat) * w_hat)   (:1056-1057)
      M3 kw; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) kw.m[i][j] = k * w_hat.m[i][j];
      M3 w2 = mul(kw, w_hat);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A_inv.m[i][j] = (((i == j) ? 1.0 : 0.0) - w_hat.m[i][j] / 2) + w2.m[i][j];
    }
    V3 ups = mul(A_inv, t);
    xi[0] = ups[0]; xi[1] = ups[1]; xi[2] = ups[2]; xi[3] = w[0]; xi[4] = w[1]; xi[5] = w[2];
  }

  // :232-244
  void predict_pose(double time_to_predict) {
    predicted_time = time_to_predict;
    double delta[6], delta_hat[6];
    logarithm_map(mul(inverse4(previous_pose), current_pose), delta);
    for (int i = 0; i < 6; ++i) delta_hat[i] = delta[i] / (current_time - previous_time) * (predicted_time - current_time);
    predicted_pose = mul(current_pose, exponential_map(delta_hat));
  }

  // :270-276
  void predict_marker_positions_in_image() {
    predicted_pixel_positions.resize(object_points.size());
    for (size_t i = 0; i < object_points.size(); ++i) predicted_pixel_positions[i] = project2d(object_points[i], predicted_pose);
  }

  // :932-960 (Eade A.14)
  static void compute_jacobian(const M4& T, const V4& wp, double fx, double fy, double J[2][6]) {
    V4 pc = mul(T, wp);
    double x = pc[0], y = pc[1], z = pc[2], z_2 = z * z;
    J[0][0] = 1 / z * fx; J[0][1] = 0; J[0][2] = -x / z_2 * fx; J[0][3] = -x * y / z_2 * fx; J[0][4] = (1 + (x * x / z_2)) * fx; J[0][5] = -y / z * fx;
    J[1][0] = 0; J[1][1] = 1 / z * fy; J[1][2] = -y / z_2 * fy; J[1][3] = -(1 + y * y / z_2) * fy; J[1][4] = x * y / z_2 * fy; J[1][5] = x / z * fy;
  }

  // :733-792
  void optimise_pose() {
    const double converged = 1e-13; const unsigned max_itr = 500;
    double A[36], b[6], dT[6];
    double fx = K[0][0], fy = K[1][1];
    std::memset(A, 0, sizeof(A));
    last_gn_iterations = 0;
    for (unsigned i = 0; i < max_itr; ++i) {
      std::memset(A, 0, sizeof(A)); std::memset(b, 0, sizeof(b));
      for (size_t j = 0; j < correspondences.size(); ++j) {
        if (correspondences[j][1] == 0) continue;
        const V4& op = object_points[correspondences[j][0] - 1];
        V2 p = project2d(op, predicted_pose);
        const V2& ip = image_points[correspondences[j][1] - 1];
        double e[2] = {ip[0] - p[0], ip[1] - p[1]};
        double J[2][6];
        compute_jacobian(predicted_pose, op, fx, fy, J);
        for (int r = 0; r < 6; ++r) {
          for (int c = 0; c < 6; ++c) A[r * 6 + c] += J[0][r] * J[0][c] + J[1][r] * J[1][c];
          b[r] += J[0][r] * e[0] + J[1][r] * e[1];
        }
      }
      ldlt_solve6(A, b, dT);
      predicted_pose = mul(exponential_map(dT), predicted_pose);
      ++last_gn_iterations;
      double mx = -1;  // norm_max :1073-1085
      for (int k = 0; k < 6; ++k) { double a = std::fabs(dT[k]); if (a > mx) mx = a; }
      if (mx <= converged) break;
    }
    inverse6(A, pose_covariance);
  }

  void update_pose() {  // :794-800
    previous_pose = current_pose; current_pose = predicted_pose;
    previous_time = current_time; current_time = predicted_time;
  }
  void optimise_and_update_pose() {  // :802-812
    optimise_pose();
    if (it_since_initialized < 2) it_since_initialized++;
    update_pose();
  }
};

// L/src/led_detector.cpp:181-224 (reads D[4] unconditionally; caller must pass >= 5 coefficients)
void distort_point(const double K[3][3], const double* D, float sx, float sy, float& ox, float& oy) {
  double fx = K[0][0], fy = K[1][1], cx = K[0][2], cy = K[1][2];
  double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
  double px = sx, py = sy;  // cv::Point2d built from Point2f
  double x = (px - cx) / fx, y = (py - cy) / fy;
  double r2 = x * x + y * y;
  double xc = x * (1. + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2);
  double yc = y * (1. + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2);
  xc = xc + (2. * p1 * x * y + p2 * (r2 + 2. * x * x));
  yc = yc + (p1 * (r2 + 2. * y * y) + 2. * p2 * x * y);
  xc = xc * fx + cx; yc = yc * fy + cy;
  ox = (float)xc; oy = (float)yc;  // push_back(cv::Point2d) into vector<Point2f>
}

// L/src/led_detector.cpp:114-179
void determine_roi(const std::vector<V2>& px, int img_w, int img_h, int border, const double K[3][3], const double* D, int roi[4]) {
  double x_min = INFINITY, x_max = 0, y_min = INFINITY, y_max = 0;
  for (const auto& p : px) {
    if (p[0] < x_min) x_min = p[0];
    if (p[0] > x_max) x_max = p[0];
    if (p[1] < y_min) y_min = p[1];
    if (p[1] > y_max) y_max = p[1];
  }
  float ax = (float)x_min, ay = (float)y_min, bx = (float)x_max, by = (float)y_max;  // Point2f corners :144-145
  float dax, day, dbx, dby;
  distort_point(K, D, ax, ay, dax, day);
  distort_point(K, D, bx, by, dbx, dby);
  double x_min_d = dax, y_min_d = day, x_max_d = dbx, y_max_d = dby;
  double x0 = std::max(0.0, std::min((double)img_w, x_min_d - border));
  double x1 = std::max(0.0, std::min((double)img_w, x_max_d + border));
  double y0 = std::max(0.0, std::min((double)img_h, y_min_d - border));
  double y1 = std::max(0.0, std::min((double)img_h, y_max_d + border));
  if (x1 - x0 < 1 || y1 - y0 < 1) { roi[0] = 0; roi[1] = 0; roi[2] = img_w; roi[3] = img_h; }
  else { roi[0] = (int)x0; roi[1] = (int)y0; roi[2] = (int)(x1 - x0); roi[3] = (int)(y1 - y0); }
}

void m4_to_rowmajor(const M4& T, double out[16]) { std::memcpy(out, &T.m[0][0], 16 * sizeof(double)); }
M4 m4_from_rowmajor(const double in[16]) { M4 T; std::memcpy(&T.m[0][0], in, 16 * sizeof(double)); return T; }

}  // namespace

// ---------------------------------------------------------------- C API (ctypes)
extern "C" {

int mpeo_solve_quartic(const double factors[5], double roots[4]) { solve_quartic(factors, roots); return 0; }

// f, P: 3 columns each, stored column after column (f[3*k + r]).  sol: 4 x (3x4 row-major).
int mpeo_p3p(const double f[9], const double P[9], double sol[48]) {
  V3 fv[3], pw[3];
  for (int k = 0; k < 3; ++k) { fv[k] = {f[3 * k], f[3 * k + 1], f[3 * k + 2]}; pw[k] = {P[3 * k], P[3 * k + 1], P[3 * k + 2]}; }
  M34 s[4];
  for (int i = 0; i < 48; ++i) sol[i] = 0;
  int rc = p3p_compute_poses(fv, pw, s);
  if (rc == 0) std::memcpy(sol, s, sizeof(s));
  return rc;
}

void* mpeo_create() { return new Oracle(); }
void mpeo_destroy(void* h) { delete (Oracle*)h; }

void mpeo_set_camera(void* h, const double K[9], const double* D, int nD) {
  Oracle* o = (Oracle*)h;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o->K[i][j] = K[3 * i + j];
  o->D.assign(D, D + nD);
}
void mpeo_set_markers(void* h, const double* xyz, int n) { ((Oracle*)h)->set_marker_positions(xyz, (unsigned)n); }
void mpeo_set_params(void* h, double back_proj_tol, double nn_tol, double certainty_thr, double valid_corr_thr) {
  Oracle* o = (Oracle*)h;
  o->back_projection_pixel_tolerance = back_proj_tol; o->nearest_neighbour_pixel_tolerance = nn_tol;
  o->certainty_threshold = certainty_thr; o->valid_correspondence_threshold = valid_corr_thr;
}
void mpeo_set_histogram_threshold(void* h, unsigned t) { ((Oracle*)h)->histogram_threshold = t; }
unsigned mpeo_get_histogram_threshold(void* h) { return ((Oracle*)h)->histogram_threshold; }
void mpeo_set_image_points(void* h, const double* pts, int n) { ((Oracle*)h)->set_image_points(pts, (unsigned)n); }
int mpeo_get_image_vectors(void* h, double* out) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->image_vectors.size(); ++i) for (int k = 0; k < 3; ++k) out[3 * i + k] = o->image_vectors[i][k];
  return (int)o->image_vectors.size();
}
unsigned mpeo_initialise(void* h) { return ((Oracle*)h)->initialise(); }
int mpeo_get_histogram(void* h, unsigned* out) {
  Oracle* o = (Oracle*)h;
  std::copy(o->last_hist.begin(), o->last_hist.end(), out);
  return (int)o->last_hist.size();
}
void mpeo_get_counters(void* h, unsigned long long out[3]) { Oracle* o = (Oracle*)h; out[0] = o->n_p3p; out[1] = o->n_finite; out[2] = o->n_voting; }
int mpeo_get_correspondences(void* h, unsigned* out) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->correspondences.size(); ++i) { out[2 * i] = o->correspondences[i][0]; out[2 * i + 1] = o->correspondences[i][1]; }
  return (int)o->correspondences.size();
}
void mpeo_set_correspondences(void* h, const unsigned* c, int n) {
  Oracle* o = (Oracle*)h; o->correspondences.resize(n);
  for (int i = 0; i < n; ++i) o->correspondences[i] = {c[2 * i], c[2 * i + 1]};
}
unsigned mpeo_check_correspondences(void* h) { return ((Oracle*)h)->check_correspondences(); }
void mpeo_find_correspondences(void* h) { ((Oracle*)h)->find_correspondences(); }
int mpeo_optimise_pose(void* h) { Oracle* o = (Oracle*)h; o->optimise_pose(); return (int)o->last_gn_iterations; }
void mpeo_optimise_and_update_pose(void* h) { ((Oracle*)h)->optimise_and_update_pose(); }
int mpeo_last_gn_iterations(void* h) { return (int)((Oracle*)h)->last_gn_iterations; }
void mpeo_update_pose(void* h) { ((Oracle*)h)->update_pose(); }
void mpeo_get_predicted_pose(void* h, double out[16]) { m4_to_rowmajor(((Oracle*)h)->predicted_pose, out); }
void mpeo_set_predicted_pose(void* h, const double in[16], double time) { Oracle* o = (Oracle*)h; o->predicted_pose = m4_from_rowmajor(in); o->predicted_time = time; }
void mpeo_get_current_pose(void* h, double out[16]) { m4_to_rowmajor(((Oracle*)h)->current_pose, out); }
void mpeo_get_previous_pose(void* h, double out[16]) { m4_to_rowmajor(((Oracle*)h)->previous_pose, out); }
void mpeo_set_state(void* h, const double cur[16], const double prev[16], double cur_t, double prev_t, unsigned it_since_init) {
  Oracle* o = (Oracle*)h; o->current_pose = m4_from_rowmajor(cur); o->previous_pose = m4_from_rowmajor(prev);
  o->current_time = cur_t; o->previous_time = prev_t; o->it_since_initialized = it_since_init;
}
void mpeo_get_covariance(void* h, double out[36]) { std::memcpy(out, ((Oracle*)h)->pose_covariance, 36 * sizeof(double)); }
void mpeo_set_predicted_time(void* h, double t) { ((Oracle*)h)->predicted_time = t; }
double mpeo_get_predicted_time(void* h) { return ((Oracle*)h)->predicted_time; }
unsigned mpeo_it_since_initialized(void* h) { return ((Oracle*)h)->it_since_initialized; }
void mpeo_predict_pose(void* h, double t) { ((Oracle*)h)->predict_pose(t); }
void mpeo_predict_marker_positions(void* h) { ((Oracle*)h)->predict_marker_positions_in_image(); }
int mpeo_get_predicted_pixels(void* h, double* out) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->predicted_pixel_positions.size(); ++i) { out[2 * i] = o->predicted_pixel_positions[i][0]; out[2 * i + 1] = o->predicted_pixel_positions[i][1]; }
  return (int)o->predicted_pixel_positions.size();
}
void mpeo_set_predicted_pixels(void* h, const double* p, int n) {
  Oracle* o = (Oracle*)h; o->predicted_pixel_positions.resize(n);
  for (int i = 0; i < n; ++i) o->predicted_pixel_positions[i] = {p[2 * i], p[2 * i + 1]};
}
// LEDDetector::determineROI on the estimator's predicted pixels; roi = x, y, w, h
void mpeo_determine_roi(void* h, int img_w, int img_h, int border, int roi[4]) {
  Oracle* o = (Oracle*)h;
  determine_roi(o->predicted_pixel_positions, img_w, img_h, border, o->K, o->D.data(), roi);
}
void mpeo_project2d(void* h, const double point[4], const double T[16], double out[2]) {
  Oracle* o = (Oracle*)h; V2 r = o->project2d({point[0], point[1], point[2], point[3]}, m4_from_rowmajor(T)); out[0] = r[0]; out[1] = r[1];
}
void mpeo_exponential_map(const double twist[6], double out[16]) { m4_to_rowmajor(Oracle::exponential_map(twist), out); }
void mpeo_logarithm_map(const double T[16], double xi[6]) { Oracle::logarithm_map(m4_from_rowmajor(T), xi); }
void mpeo_distort_point(void* h, float x, float y, float out[2]) { Oracle* o = (Oracle*)h; distort_point(o->K, o->D.data(), x, y, out[0], out[1]); }

// Batch helper for CPU baseline timing: runs the cold pose path (setImagePoints -> initialise ->
// optimiseAndUpdatePose) on n_frames detection sets with a fresh state each; returns #updated.
int mpeo_cold_pose_batch(void* h, const double* dets, const int* n_det, int n_frames, int max_det, double* poses_out, int* updated_out) {
  Oracle* o = (Oracle*)h; int ok = 0;
  for (int f = 0; f < n_frames; ++f) {
    o->it_since_initialized = 0;
    int upd = 0;
    if (n_det[f] >= 4) {
      o->set_image_points(dets + (size_t)f * max_det * 2, (unsigned)n_det[f]);
      if (o->initialise() == 1) { o->optimise_and_update_pose(); upd = 1; }
    }
    if (updated_out) updated_out[f] = upd;
    if (poses_out) m4_to_rowmajor(o->predicted_pose, poses_out + 16 * (size_t)f);
    ok += upd;
  }
  return ok;
}

}  // extern "C"
