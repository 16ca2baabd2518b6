// TEST INFRASTRUCTURE — host build of the K2 tier-1 pre-test (rpg_monocular_pose_estimator_b200/csrc/p3p_tier1.cuh) next to
// the host build of the product's exact P3P path (csrc/p3p_device.cuh, whose arithmetic the GPU tests show to be the
// oracle's), to measure on the CPU what the conservativeness argument rests on:
//   * violations: problems that vote under the exact arithmetic but that tier 1 would have rejected     (must be 0)
//   * the closest call: the smallest exact distance (minus the tolerance) among rejected problems       (must be >= 0; the
//     margin makes it >= margin minus the deviation below)
//   * the deviation between tier-1 and exact roots / back-projections on problems tier 1 does not flag
//   * how many problems survive (cost model of the two-tier sweep).
// Built by oracle/Makefile into oracle/libtier1_check.so; loaded by tests/test_cpu_k2_tier1.py only.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include "../rpg_monocular_pose_estimator_b200/csrc/p3p_device.cuh"
#include "../rpg_monocular_pose_estimator_b200/csrc/p3p_tier1.cuh"

using namespace mpe;
#ifdef MPE_T1_DEBUG
static FILE* g_dump = nullptr;
extern "C" void t1c_dump_to(const char* path) { if (g_dump) fclose(g_dump); g_dump = path ? fopen(path, "wb") : nullptr; }
#endif

namespace {
constexpr double kCondMin = 1e-6;   // k2_p3p_sweep.cu

int nth_unused(int m, int a, int b, int c) { m += (m >= a); m += (m >= b); m += (m >= c); return m; }

v3 bearing(const double K[9], double u, double v) {
  double x = (u - K[2]) / K[0], y = (v - K[5]) / K[4], z = 1, n = std::sqrt(x * x + y * y + z * z);
  return v_make(x / n, y / n, z / n);
}
}  // namespace

extern "C" {

struct t1c_stats {
  long long problems, conditioned, survivors, flagged, voting_problems, violations;
  double closest_call;        // min over rejected problems of (exact nearest distance - tolerance)
  double max_root_dev;        // max |rho_t1 - rho_exact| over finite exact hypotheses of non-flagged problems
  double max_pixel_dev;       // max |pixel_t1 - pixel_exact| over their back-projections that land in or near the image
  long long compared;
};

// One frame: K row-major, markers n_obj x 3, det n_det x 2 (undistorted pixels), tolerance and margin in pixels.
static int g_fp32 = 1;
extern "C" void t1c_set_fp32(int on) { g_fp32 = on; }

void t1c_frame(const double K[9], const double* mk, int n_obj, const double* det, int n_det, double tol, double margin, t1c_stats* S) {
  const double r = tol + margin;
  const double tol_sq = tol * tol;
  std::vector<v3> bear(n_det);
  for (int i = 0; i < n_det; ++i) bear[i] = bearing(K, det[2 * i], det[2 * i + 1]);
  for (int d0 = 0; d0 < n_det; ++d0) for (int d1 = d0 + 1; d1 < n_det; ++d1) for (int d2 = d1 + 1; d2 < n_det; ++d2) {
    P3PCamera Cm;
    p3p_camera_frame(bear[d0], bear[d1], bear[d2], Cm);
    double Mc[9];
    const v3 e[3] = {Cm.e1, Cm.e2, Cm.e3};
    for (int rr = 0; rr < 3; ++rr) for (int q = 0; q < 3; ++q) Mc[3 * rr + q] = K[3 * rr] * e[q].x + K[3 * rr + 1] * e[q].y + K[3 * rr + 2] * e[q].z;
    const bool cam_ok = Cm.sin12 > kCondMin;
    for (int o0 = 0; o0 < n_obj; ++o0) for (int o1 = 0; o1 < n_obj; ++o1) for (int o2 = 0; o2 < n_obj; ++o2) {
      if (o0 == o1 || o0 == o2 || o1 == o2) continue;
      auto P = [&](int i) { return v_make(mk[3 * i], mk[3 * i + 1], mk[3 * i + 2]); };
      P3PWorld W0;
      p3p_world_frame(P(o0), P(o1), P(o2), W0);
      if (W0.cross_norm == 0.0) continue;                          // colinear: no hypotheses (p3p.cpp:77-80)
      ++S->problems;
      P3PWorld W;
      if (Cm.swap) p3p_world_frame(P(o1), P(o0), P(o2), W); else W = W0;
      const double len13 = std::sqrt(W.p_1 * W.p_1 + W.p_2 * W.p_2);
      const bool world_ok = W.cross_norm > kCondMin * W.d_12 * len13;
      const int oa = std::min(o0, std::min(o1, o2)), oc = std::max(o0, std::max(o1, o2)), ob = o0 + o1 + o2 - oa - oc;
      const int nu_obj = n_obj - 3, nu_det = n_det - 3;

      // ---- exact path (host build of the product's arithmetic)
      P3PSetup St;
      p3p_assemble(Cm, W, St);
      bool votes = false;
      double exact_min = HUGE_VAL;
      double ex_rho[4]; bool ex_fin[4]; double ex_px[4][16][2];
      for (int k = 0; k < 4; ++k) {
        double H[12];
        ex_fin[k] = false; ex_rho[k] = St.roots[k];
        if (!p3p_solution(St, k, H)) continue;
        if (!h_is_finite(H)) continue;
        ex_fin[k] = true;
        double Hi[12], KT[12];
        h_inverse(H, Hi); kt_product(K, Hi, KT);
        for (int m = 0; m < nu_obj; ++m) { const int ll = nth_unused(m, oa, ob, oc); kt_project(KT, mk[3 * ll], mk[3 * ll + 1], mk[3 * ll + 2], ex_px[k][m][0], ex_px[k][m][1]); }
        for (int i = 0; i < nu_det; ++i) {
          const int kk = nth_unused(i, d0, d1, d2);
          double best = HUGE_VAL;
          for (int m = 0; m < nu_obj; ++m) {
            const double dx = det[2 * kk] - ex_px[k][m][0], dy = det[2 * kk + 1] - ex_px[k][m][1], dd = dx * dx + dy * dy;
            if (dd < best) best = dd;
          }
          if (std::sqrt(best) < tol) votes = true;                 // pose_estimator.cpp:671
          (void)tol_sq;
          if (best == best) exact_min = std::min(exact_min, std::sqrt(best));
        }
      }
      if (votes) ++S->voting_problems;

      // ---- tier 1
      bool maybe = true, flagged = false;
      if (cam_ok && world_ok) {
        ++S->conditioned;
        T1Roots R;
        t1_quartic_roots(Cm.f_1, Cm.f_2, Cm.b, W.p_1, W.p_2, W.d_12, R);
        if (R.maybe) { flagged = true; }
        else {
          maybe = false;
          T1Pose Pk[4]; int st[4];
          for (int k = 0; k < 4; ++k) {
            st[k] = t1_pose(R.rho[k], Cm.f_1, Cm.f_2, Cm.b, W.p_1, W.p_2, W.d_12, Pk[k]);
            if (st[k] == 2) { flagged = true; maybe = true; }
          }
          double t1px[4][16][3];
          for (int k = 0; k < 4 && !flagged; ++k) {
            if (st[k] != 1) continue;
            for (int m = 0; m < nu_obj; ++m) {
              const int ll = nth_unused(m, oa, ob, oc);
              const v3 dx = v_sub(P(ll), W.P1);
              double au, av, az, l1;
              if (g_fp32) {
                float Mcf[9]; for (int e = 0; e < 9; ++e) Mcf[e] = (float)Mc[e];
                float fu, fv, fz, fl;
                t1_project_f(t1_pose_f(Pk[k]), Mcf, (float)v_dot(W.n1, dx), (float)v_dot(W.n2, dx), (float)v_dot(W.n3, dx), fu, fv, fz, fl);
                t1px[k][m][0] = fu; t1px[k][m][1] = fv; t1px[k][m][2] = fz;
                if (!(std::fabs(fz) >= 1e-3f * fl) || !(fl >= 1e-3f * (float)W.d_12)) maybe = true;
                const float rf = (float)r, lim = rf * rf * (fz * fz);
                for (int i = 0; i < nu_det; ++i) {
                  const int kk = nth_unused(i, d0, d1, d2);
                  const float eu = fu - (float)det[2 * kk] * fz, ev = fv - (float)det[2 * kk + 1] * fz;
                  if (!(eu * eu + ev * ev > lim)) maybe = true;
                }
                continue;
              }
              t1_project(Pk[k], Mc, v_dot(W.n1, dx), v_dot(W.n2, dx), v_dot(W.n3, dx), au, av, az, l1);
              t1px[k][m][0] = au; t1px[k][m][1] = av; t1px[k][m][2] = az;
              if (!(std::fabs(az) >= 1e-3 * l1) || !(l1 >= 1e-3 * W.d_12)) maybe = true;
              const double lim = r * r * (az * az);
              for (int i = 0; i < nu_det; ++i) {
                const int kk = nth_unused(i, d0, d1, d2);
                const double eu = au - det[2 * kk] * az, ev = av - det[2 * kk + 1] * az;
                if (!(eu * eu + ev * ev > lim)) maybe = true;
              }
            }
          }
          // deviation statistics: every finite exact hypothesis against the nearest tier-1 root
          if (!flagged) {
            for (int k = 0; k < 4; ++k) {
              if (!ex_fin[k]) continue;
              int best = -1; double bd = HUGE_VAL;
              for (int q = 0; q < 4; ++q) if (st[q] == 1 && std::fabs(R.rho[q] - ex_rho[k]) < bd) { bd = std::fabs(R.rho[q] - ex_rho[k]); best = q; }
              if (best < 0) { S->max_root_dev = std::max(S->max_root_dev, 1.0); continue; }   // an exact finite hypothesis without tier-1 partner
#ifdef MPE_T1_DEBUG
              if (g_dump) { double row[8] = {bd, g_t1_debug.al, g_t1_debug.be, g_t1_debug.ga, g_t1_debug.z, g_t1_debug.kappa, g_t1_debug.d1, g_t1_debug.d2}; fwrite(row, 8, 8, g_dump); }
#endif
              if (bd > 1e-6 && getenv("T1C_DEBUG")) {
                fprintf(stderr, "dev %.3e  exact roots %.12f %.12f %.12f %.12f  t1 roots %.12f %.12f %.12f %.12f  f1 %.6f f2 %.6f b %.6f p1 %.6f p2 %.6f d12 %.6f\n", bd,
                        ex_rho[0], ex_rho[1], ex_rho[2], ex_rho[3], R.rho[0], R.rho[1], R.rho[2], R.rho[3], Cm.f_1, Cm.f_2, Cm.b, W.p_1, W.p_2, W.d_12);
              }
              S->max_root_dev = std::max(S->max_root_dev, bd);
              ++S->compared;
              for (int m = 0; m < nu_obj; ++m) {
                const double u = t1px[best][m][0] / t1px[best][m][2], v = t1px[best][m][1] / t1px[best][m][2];
                if (ex_px[k][m][0] > -200 && ex_px[k][m][0] < 2200 && ex_px[k][m][1] > -200 && ex_px[k][m][1] < 1300)   // where a detection can be
                  S->max_pixel_dev = std::max(S->max_pixel_dev, std::max(std::fabs(u - ex_px[k][m][0]), std::fabs(v - ex_px[k][m][1])));
              }
            }
          }
        }
      }
      if (flagged) ++S->flagged;
      if (maybe) ++S->survivors;
      else {
        if (votes) ++S->violations;
        S->closest_call = std::min(S->closest_call, exact_min - tol);
      }
    }
  }
}

void t1c_init(t1c_stats* S) { *S = t1c_stats(); S->closest_call = HUGE_VAL; }

}  // extern "C"
