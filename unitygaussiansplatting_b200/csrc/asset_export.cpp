// asset_export.cpp -- "bake transform" of exported splats (libgsplat_asset.so): the _ExportTransformFlags != 0 branch of
// CSExportData (S/SplatUtilities.compute:626-643) as a host pass over the raw .ply attribute records that
// gs_export_splats returns.  Position by the object-to-world matrix, orientation by the transform's rotation (with the
// reference's axis-flip rule for negative scale), log-scale by |scale|, and the SH bands 1..3 rotated into world space.
//
// The SH rotation is NOT the reference's closed form (S/SphericalHarmonics.hlsl, after andrewwillmott/sh-lib): the rotation
// is the same for every splat of an export, so the three band matrices (3x3, 5x5, 7x7) are solved once, in float64, by
// least squares over sample directions -- c' = A^+ A_R c with A[s][j] = Y_j(d_s), A_R[s][k] = Y_k(R^T d_s) -- using the
// real-SH basis ShadeSH evaluates (S/GaussianSplatting.hlsl:139-179).  Then every splat costs three small mat-vecs per
// colour channel.  Equal to the closed form up to float rounding; tests/test_export.py checks the property that defines
// it: a baked export renders, untransformed, like the original asset under the transform.
#include "../../include/gsplat_asset.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

int gsa_usable_threads();  // asset_cluster_bc7.cpp

namespace {

// real SH basis of bands 1..3 in ShadeSH's order and sign convention (15 functions)
void sh_basis(const double d[3], double Y[15]) {
  const double x = d[0], y = d[1], z = d[2];
  const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  Y[0] = -0.4886025119029199 * y; Y[1] = 0.4886025119029199 * z; Y[2] = -0.4886025119029199 * x;
  Y[3] = 1.0925484305920792 * xy; Y[4] = -1.0925484305920792 * yz; Y[5] = 0.31539156525252005 * (2 * zz - xx - yy);
  Y[6] = -1.0925484305920792 * xz; Y[7] = 0.5462742152960396 * (xx - yy);
  Y[8] = -0.5900435899266435 * y * (3 * xx - yy); Y[9] = 2.890611442640554 * xy * z; Y[10] = -0.4570457994644658 * y * (4 * zz - xx - yy);
  Y[11] = 0.3731763325901154 * z * (2 * zz - 3 * xx - 3 * yy); Y[12] = -0.4570457994644658 * x * (4 * zz - xx - yy);
  Y[13] = 1.445305721320277 * z * (xx - yy); Y[14] = -0.5900435899266435 * x * (xx - 3 * yy);
}

// solves (A^T A) X = A^T B for one band: A, B are S x n (row-major), X is n x n
bool band_matrix(int n, int S, const std::vector<double> &A, const std::vector<double> &B, double *X) {
  std::vector<double> M((size_t)n * n, 0.0), R((size_t)n * n, 0.0);
  for (int s = 0; s < S; ++s)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        M[i * n + j] += A[(size_t)s * n + i] * A[(size_t)s * n + j];
        R[i * n + j] += A[(size_t)s * n + i] * B[(size_t)s * n + j];
      }
  for (int c = 0; c < n; ++c) {  // Gauss-Jordan with partial pivoting on [M | R]
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (std::fabs(M[r * n + c]) > std::fabs(M[piv * n + c])) piv = r;
    if (std::fabs(M[piv * n + c]) < 1e-12) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) { std::swap(M[c * n + j], M[piv * n + j]); std::swap(R[c * n + j], R[piv * n + j]); }
    const double inv = 1.0 / M[c * n + c];
    for (int j = 0; j < n; ++j) { M[c * n + j] *= inv; R[c * n + j] *= inv; }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = M[r * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) { M[r * n + j] -= f * M[c * n + j]; R[r * n + j] -= f * R[c * n + j]; }
    }
  }
  std::memcpy(X, R.data(), sizeof(double) * n * n);
  return true;
}

}  // namespace

extern "C" {

// records: n x 62 raw attribute values, modified in place.  o2w: column-major 4x4 (tr.localToWorldMatrix);
// rot_xyzw / scale: tr.localRotation / tr.localScale (R/GaussianSplatRenderer.cs:941-953).
int gsa_bake_transform(float *records, uint32_t n, const float o2w[16], const float rot_xyzw[4], const float scale[3]) {
  if ((!records && n) || !o2w || !rot_xyzw || !scale) return -1;
  // CalcSHRotMatrix (S/SplatUtilities.compute:589-608): the rows of the upper 3x3, each normalised
  double R[3][3];
  for (int r = 0; r < 3; ++r) {
    const double a = o2w[0 * 4 + r], b = o2w[1 * 4 + r], c = o2w[2 * 4 + r];
    const double len = std::sqrt(a * a + b * b + c * c);
    if (!(len > 0.0)) return -1;
    R[r][0] = a / len; R[r][1] = b / len; R[r][2] = c / len;
  }
  // band matrices from 64 Fibonacci-sphere directions: world-space lobe f_w(d) = f_o(R^T d)
  const int S = 64;
  std::vector<double> A1(S * 3), B1(S * 3), A2(S * 5), B2(S * 5), A3(S * 7), B3(S * 7);
  for (int s = 0; s < S; ++s) {
    const double zc = 1.0 - 2.0 * (s + 0.5) / S, rad = std::sqrt(std::fmax(0.0, 1.0 - zc * zc)), phi = s * 2.399963229728653;
    const double d[3] = {rad * std::cos(phi), rad * std::sin(phi), zc};
    const double dr[3] = {R[0][0] * d[0] + R[1][0] * d[1] + R[2][0] * d[2], R[0][1] * d[0] + R[1][1] * d[1] + R[2][1] * d[2],
                          R[0][2] * d[0] + R[1][2] * d[1] + R[2][2] * d[2]};  // R^T d
    double Y[15], Yr[15];
    sh_basis(d, Y);
    sh_basis(dr, Yr);
    for (int k = 0; k < 3; ++k) { A1[s * 3 + k] = Y[k]; B1[s * 3 + k] = Yr[k]; }
    for (int k = 0; k < 5; ++k) { A2[s * 5 + k] = Y[3 + k]; B2[s * 5 + k] = Yr[3 + k]; }
    for (int k = 0; k < 7; ++k) { A3[s * 7 + k] = Y[8 + k]; B3[s * 7 + k] = Yr[8 + k]; }
  }
  double M1[9], M2[25], M3[49];   // c'_j = sum_k M[j][k] c_k
  if (!band_matrix(3, S, A1, B1, M1) || !band_matrix(5, S, A2, B2, M2) || !band_matrix(7, S, A3, B3, M3)) return -1;

  const float qa[4] = {rot_xyzw[0], rot_xyzw[1], rot_xyzw[2], rot_xyzw[3]};
  const float ls[3] = {std::log(std::fabs(scale[0])), std::log(std::fabs(scale[1])), std::log(std::fabs(scale[2]))};
#pragma omp parallel for schedule(static) num_threads(gsa_usable_threads())
  for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
    float *r = records + (size_t)ii * 62;
    // position, :628
    const float px = r[0], py = r[1], pz = r[2];
    r[0] = o2w[0] * px + o2w[4] * py + o2w[8] * pz + o2w[12];
    r[1] = o2w[1] * px + o2w[5] * py + o2w[9] * pz + o2w[13];
    r[2] = o2w[2] * px + o2w[6] * py + o2w[10] * pz + o2w[14];
    // rotation: stored wxyz; axis flips for negative scale (:631-636), then QuatMul(transform, splat) (:637)
    float b[4] = {r[59], r[60], r[61], r[58]};  // xyzw
    if (scale[0] < 0) { b[1] = -b[1]; b[2] = -b[2]; }
    if (scale[1] < 0) { b[0] = -b[0]; b[2] = -b[2]; }
    if (scale[2] < 0) { b[0] = -b[0]; b[1] = -b[1]; }
    const float qx = qa[3] * b[0] + qa[0] * b[3] + qa[1] * b[2] - qa[2] * b[1];
    const float qy = qa[3] * b[1] - qa[0] * b[2] + qa[1] * b[3] + qa[2] * b[0];
    const float qz = qa[3] * b[2] + qa[0] * b[1] - qa[1] * b[0] + qa[2] * b[3];
    const float qw = qa[3] * b[3] - qa[0] * b[0] - qa[1] * b[1] - qa[2] * b[2];
    r[58] = qw; r[59] = qx; r[60] = qy; r[61] = qz;
    // scale is stored as log(scale): src.scale *= abs(scale), :638
    r[55] += ls[0]; r[56] += ls[1]; r[57] += ls[2];
    // SH bands 1..3, channel-major f_rest (15 per channel); band 0 (f_dc) is rotation invariant
    for (int ch = 0; ch < 3; ++ch) {
      float *c = r + 9 + ch * 15;
      double in[15], out[15];
      for (int k = 0; k < 15; ++k) in[k] = c[k];
      for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += M1[j * 3 + k] * in[k]; out[j] = a; }
      for (int j = 0; j < 5; ++j) { double a = 0; for (int k = 0; k < 5; ++k) a += M2[j * 5 + k] * in[3 + k]; out[3 + j] = a; }
      for (int j = 0; j < 7; ++j) { double a = 0; for (int k = 0; k < 7; ++k) a += M3[j * 7 + k] * in[8 + k]; out[8 + j] = a; }
      for (int k = 0; k < 15; ++k) c[k] = (float)out[k];
    }
  }
  return 0;
}

}  // extern "C"
