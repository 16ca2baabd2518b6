#include <cstdio>
#include <cuda_runtime.h>
__device__ void ldlt_solve6(const double Ain[36], const double bin[6], double x[6]) {
  double A[6][6];
  int perm[6];
  for (int i = 0; i < 6; ++i) { perm[i] = i; for (int j = 0; j < 6; ++j) A[i][j] = Ain[i * 6 + j]; }
  for (int k = 0; k < 6; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
    for (int i = k + 1; i < 6; ++i) if (fabs(A[i][i]) > best) { best = fabs(A[i][i]); p = i; }
    if (p != k) {
      for (int j = 0; j < 6; ++j) { double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
      for (int i = 0; i < 6; ++i) { double t = A[i][k]; A[i][k] = A[i][p]; A[i][p] = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    double d = A[k][k];
    for (int i = k + 1; i < 6; ++i)
      for (int j = k + 1; j <= i; ++j) A[i][j] -= A[i][k] * A[j][k] / d;
    for (int i = k + 1; i < 6; ++i) A[i][k] /= d;
    for (int i = k + 1; i < 6; ++i) for (int j = k + 1; j < i; ++j) A[j][i] = A[i][j];
  }
  double y[6];
  for (int i = 0; i < 6; ++i) y[i] = bin[perm[i]];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
  for (int i = 0; i < 6; ++i) y[i] /= A[i][i];
  for (int i = 5; i >= 0; --i) for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
  for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}
__device__ void inverse6(const double Ain[36], double out[36]) {
  double a[6][12];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { a[i][j] = Ain[i * 6 + j]; a[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  for (int k = 0; k < 6; ++k) {
    int p = k;
    double best = fabs(a[k][k]);
    for (int i = k + 1; i < 6; ++i) if (fabs(a[i][k]) > best) { best = fabs(a[i][k]); p = i; }
    if (p != k) for (int j = 0; j < 12; ++j) { double t = a[k][j]; a[k][j] = a[p][j]; a[p][j] = t; }
    double piv = a[k][k];
    for (int j = 0; j < 12; ++j) a[k][j] /= piv;
    for (int i = 0; i < 6; ++i) if (i != k) {
      double fct = a[i][k];
      if (fct != 0) for (int j = 0; j < 12; ++j) a[i][j] -= fct * a[k][j];
    }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) out[i * 6 + j] = a[i][6 + j];
}
__device__ __noinline__ void ldlt_noinline(const double Ain[36], const double bin[6], double x[6]) {
  double A[6][6];
  int perm[6];
  for (int i = 0; i < 6; ++i) { perm[i] = i; for (int j = 0; j < 6; ++j) A[i][j] = Ain[i * 6 + j]; }
  for (int k = 0; k < 6; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
    for (int i = k + 1; i < 6; ++i) if (fabs(A[i][i]) > best) { best = fabs(A[i][i]); p = i; }
    if (p != k) {
      for (int j = 0; j < 6; ++j) { double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
      for (int i = 0; i < 6; ++i) { double t = A[i][k]; A[i][k] = A[i][p]; A[i][p] = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    double d = A[k][k];
    for (int i = k + 1; i < 6; ++i)
      for (int j = k + 1; j <= i; ++j) A[i][j] -= A[i][k] * A[j][k] / d;
    for (int i = k + 1; i < 6; ++i) A[i][k] /= d;
    for (int i = k + 1; i < 6; ++i) for (int j = k + 1; j < i; ++j) A[j][i] = A[i][j];
  }
  double y[6];
  for (int i = 0; i < 6; ++i) y[i] = bin[perm[i]];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
  for (int i = 0; i < 6; ++i) y[i] /= A[i][i];
  for (int i = 5; i >= 0; --i) for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
  for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

// variant: no data swaps, permutation by indirection
__device__ void ldlt_indirect(const double Ain[36], const double bin[6], double x[6]) {
  double A[6][6]; int pm[6];
  for (int i = 0; i < 6; ++i) { pm[i] = i; for (int j = 0; j < 6; ++j) A[i][j] = Ain[i*6+j]; }
  double L[6][6]; double D[6];
  for (int k = 0; k < 6; ++k) {
    int p = k; double best = fabs(A[pm[k]][pm[k]]);
    for (int i = k+1; i < 6; ++i) { double v = fabs(A[pm[i]][pm[i]]); if (v > best) { best = v; p = i; } }
    int t = pm[k]; pm[k] = pm[p]; pm[p] = t;
    if (p != k) for (int j = 0; j < k; ++j) { double tt = L[k][j]; L[k][j] = L[p][j]; L[p][j] = tt; }
    double d = A[pm[k]][pm[k]]; D[k] = d;
    for (int i = k+1; i < 6; ++i) for (int j = k+1; j <= i; ++j) { double v = A[pm[i]][pm[j]] - A[pm[i]][pm[k]] * A[pm[j]][pm[k]] / d; A[pm[i]][pm[j]] = v; A[pm[j]][pm[i]] = v; }
    for (int i = k+1; i < 6; ++i) L[i][k] = A[pm[i]][pm[k]] / d;
  }
  double y[6];
  for (int i = 0; i < 6; ++i) y[i] = bin[pm[i]];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) y[i] -= L[i][j] * y[j];
  for (int i = 0; i < 6; ++i) y[i] /= D[i];
  for (int i = 5; i >= 0; --i) for (int j = i+1; j < 6; ++j) y[i] -= L[j][i] * y[j];
  for (int i = 0; i < 6; ++i) x[pm[i]] = y[i];
}
__global__ void k(const double* A, const double* b, double* out) {
  double x[6];
  ldlt_solve6(A, b, x); for (int i = 0; i < 6; ++i) out[i] = x[i];
  ldlt_noinline(A, b, x); for (int i = 0; i < 6; ++i) out[6+i] = x[i];
  ldlt_indirect(A, b, x); for (int i = 0; i < 6; ++i) out[12+i] = x[i];
  double inv[36]; inverse6(A, inv); for (int i = 0; i < 36; ++i) out[18+i] = inv[i];
}
int main() {
  double A[36] = {2.080222644e+06,0.000000000e+00,-8.772351340e+05,2.397801909e+04,2.338104438e+06,6.608531743e+04,
 0.000000000e+00,2.087715491e+06,6.789350502e+04,-1.996595328e+06,-2.406438659e+04,8.298878459e+05,
 -8.772351340e+05,6.789350502e+04,3.915441484e+05,-7.567669473e+04,-9.896135336e+05,8.636749828e+01,
 2.397801909e+04,-1.996595328e+06,-7.567669473e+04,1.920629166e+06,5.000759147e+04,-7.862923961e+05,
 2.338104438e+06,-2.406438659e+04,-9.896135336e+05,5.000759147e+04,2.634574442e+06,6.497156546e+04,
 6.608531743e+04,8.298878459e+05,8.636749828e+01,-7.862923961e+05,6.497156546e+04,3.451675316e+05};
  double b[6] = {-50.472760862,41.050268327,24.471941902,-38.19815302,-49.22946467,12.828563939};
  double *dA, *db, *dout; cudaMalloc(&dA, sizeof(A)); cudaMalloc(&db, sizeof(b)); cudaMalloc(&dout, 64*8);
  cudaMemcpy(dA, A, sizeof(A), cudaMemcpyHostToDevice); cudaMemcpy(db, b, sizeof(b), cudaMemcpyHostToDevice);
  k<<<1,1>>>(dA, db, dout); double out[64]; cudaError_t e = cudaMemcpy(out, dout, 54*8, cudaMemcpyDeviceToHost);
  printf("err %s\n", cudaGetErrorString(e));
  const char* names[3] = {"as-is", "noinline", "indirect"};
  for (int v = 0; v < 3; ++v) { printf("%-9s", names[v]); for (int i = 0; i < 6; ++i) printf(" % .6e", out[6*v+i]); printf("\n"); }
  // check inverse: A * inv
  double mx = 0; for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int q = 0; q < 6; ++q) s += A[i*6+q]*out[18+q*6+j]; double r = fabs(s - (i==j)); if (r > mx) mx = r; }
  printf("inverse6 max |A*inv - I| = %.3e\n", mx);
}
