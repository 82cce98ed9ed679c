/* shell.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): thin-shell helpers and the ANDES
 * quadrilateral STR24.  Follows src/vpmStress/elStressModule.f90:738-849 (STR22a), :1005-1078
 * (STR24), src/vpmStress/strainAndStressUtils.f90:101-481, src/Femlib/pmatStiff.f90:23-127,
 * src/Femlib/isoMatModule.f90:21-57, src/vpmUtilities/manipMatrixModule.f90:373-410 (invert33). */
#include "oracle.h"
#include <math.h>
#include <string.h>

static void cross(const double a[3], const double b[3], double c[3])
{
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3])
{
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* isoMatModule.f90:21-38; C column-major 3x3 */
void orc_iso_mat2d(double emod, double rnu, double C[9])
{
  memset(C, 0, 9 * sizeof(double));
  C[0] = emod / (1.0 - rnu * rnu); /* C(1,1) */
  C[3] = rnu * C[0];               /* C(1,2) */
  C[1] = C[3];                     /* C(2,1) */
  C[4] = C[0];                     /* C(2,2) */
  C[8] = 0.5 * emod / (1.0 + rnu); /* C(3,3) */
}

/* isoMatModule.f90:41-57 */
void orc_iso_mat2d_inv(double emod, double rnu, double C[9])
{
  memset(C, 0, 9 * sizeof(double));
  C[0] = 1.0 / emod;
  C[3] = -rnu / emod;
  C[1] = C[3];
  C[4] = C[0];
  C[8] = 2.0 * (1.0 + rnu) / emod;
}

/* manipMatrixModule.f90:373-410; a, b column-major 3x3; returns -1 if singular */
static int invert33(const double *a, double *b)
{
#define A(i, j) a[(i - 1) + 3 * (j - 1)]
#define Bm(i, j) b[(i - 1) + 3 * (j - 1)]
  double det = A(1, 1) * (A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3)) -
               A(1, 2) * (A(2, 1) * A(3, 3) - A(3, 1) * A(2, 3)) +
               A(1, 3) * (A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2));
  if (fabs(det) < ORC_EPSDIV0) {
    for (int i = 0; i < 9; i++) b[i] = ORC_HUGE;
    return -1;
  }
  Bm(1, 1) = (A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3)) / det;
  Bm(1, 2) = -(A(1, 2) * A(3, 3) - A(3, 2) * A(1, 3)) / det;
  Bm(1, 3) = (A(1, 2) * A(2, 3) - A(2, 2) * A(1, 3)) / det;
  Bm(2, 1) = -(A(2, 1) * A(3, 3) - A(3, 1) * A(2, 3)) / det;
  Bm(2, 2) = (A(1, 1) * A(3, 3) - A(3, 1) * A(1, 3)) / det;
  Bm(2, 3) = -(A(1, 1) * A(2, 3) - A(2, 1) * A(1, 3)) / det;
  Bm(3, 1) = (A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2)) / det;
  Bm(3, 2) = -(A(1, 1) * A(3, 2) - A(3, 1) * A(1, 2)) / det;
  Bm(3, 3) = (A(1, 1) * A(2, 2) - A(2, 1) * A(1, 2)) / det;
#undef A
#undef Bm
  return 0;
}

/* pmatStiff.f90:23-127.  pmat column-major (6*nnod)^2.  nnod <= 8. */
int orc_pmat_stiff(int nnod, const double *x, const double *y, const double *z, double *pmat)
{
  const int n6 = 6 * nnod;
  double coorRel[3][8], rmat[48][6], rsmat[6], sub[9], subinv[9];
  const double *coor[3] = {x, y, z};

  for (int i = 0; i < 3; i++) {
    double c = 0.0;
    for (int k = 0; k < nnod; k++) c += coor[i][k];
    c = c / (double)nnod;
    for (int k = 0; k < nnod; k++) coorRel[i][k] = coor[i][k] - c;
  }

  memset(rmat, 0, sizeof(rmat));
  for (int i = 0; i < nnod; i++) {
    int j = i * 6; /* rmat(j+k, m) -> rmat[j+k-1][m-1] */
    rmat[j + 0][0] = 1.0;
    rmat[j + 1][1] = 1.0;
    rmat[j + 2][2] = 1.0;
    rmat[j + 1][3] = -coorRel[2][i];
    rmat[j + 2][3] = coorRel[1][i];
    rmat[j + 3][3] = 1.0;
    rmat[j + 0][4] = coorRel[2][i];
    rmat[j + 2][4] = -coorRel[0][i];
    rmat[j + 4][4] = 1.0;
    rmat[j + 0][5] = -coorRel[1][i];
    rmat[j + 1][5] = coorRel[0][i];
    rmat[j + 5][5] = 1.0;
  }

  for (int i = 0; i < 6; i++) {
    double c = 0.0;
    for (int k = 0; k < n6; k++) c += rmat[k][i] * rmat[k][i];
    c = sqrt(c);
    for (int k = 0; k < n6; k++) rmat[k][i] = rmat[k][i] / c;
  }

  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < n6; k++) s += rmat[k][i + 3] * rmat[k][j + 3];
      sub[i + 3 * j] = s;
    }
  if (invert33(sub, subinv) < 0) return -1;

  for (int i = 0; i < n6; i++) {
    rsmat[0] = rmat[i][0];
    rsmat[1] = rmat[i][1];
    rsmat[2] = rmat[i][2];
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += rmat[i][3 + k] * subinv[k + 3 * j];
      rsmat[3 + j] = s;
    }
    for (int j = 0; j < n6; j++) {
      double s = 0.0;
      for (int k = 0; k < 6; k++) s += rsmat[k] * rmat[j][k];
      pmat[i + n6 * j] = -s;
    }
    pmat[i + n6 * i] = pmat[i + n6 * i] + 1.0;
  }
  return 0;
}

/* strainAndStressUtils.f90:297-336 */
static void globalized_x(const double VZ[3], double V1[3])
{
  const double somewhatSmall = 0.01;
  double V2[3], len2;
  if (fabs(VZ[1]) > somewhatSmall || fabs(VZ[2]) > somewhatSmall) {
    V1[0] = VZ[1] * VZ[1] + VZ[2] * VZ[2];
    V1[1] = -VZ[0] * VZ[1];
    V1[2] = -VZ[0] * VZ[2];
  } else {
    V2[0] = -VZ[1] * VZ[0];
    V2[1] = VZ[0] * VZ[0] + VZ[2] * VZ[2];
    V2[2] = -VZ[1] * VZ[2];
    cross(V2, VZ, V1);
  }
  len2 = dot3(V1, V1);
  if (len2 > ORC_EPSDIV0 * ORC_EPSDIV0) {
    double l = sqrt(len2);
    V1[0] /= l; V1[1] /= l; V1[2] /= l;
  } else
    V1[0] = V1[1] = V1[2] = 0.0;
}

/* strainAndStressUtils.f90:339-434 (doGlobalize absent) */
int orc_shell_element_axes(int nenod, const double *X, const double *Y, const double *Z,
                           double V1[3], double V2[3], double V3[3])
{
  double VN;
  if (nenod == 3) {
    V1[0] = X[1] - X[0]; V1[1] = Y[1] - Y[0]; V1[2] = Z[1] - Z[0];
    V2[0] = X[2] - X[0]; V2[1] = Y[2] - Y[0]; V2[2] = Z[2] - Z[0];
  } else if (nenod == 4) {
    V1[0] = X[2] - X[0]; V1[1] = Y[2] - Y[0]; V1[2] = Z[2] - Z[0];
    V2[0] = X[3] - X[1]; V2[1] = Y[3] - Y[1]; V2[2] = Z[3] - Z[1];
  } else
    return -1;
  cross(V1, V2, V3);
  VN = dot3(V3, V3);
  if (VN > ORC_EPSDIV0 * ORC_EPSDIV0) {
    double l = sqrt(VN);
    V3[0] /= l; V3[1] /= l; V3[2] /= l;
  } else
    return 2;

  if (nenod == 4) {
    V1[0] = X[1] - X[0]; V1[1] = Y[1] - Y[0]; V1[2] = Z[1] - Z[0];
    cross(V3, V1, V2);
    cross(V2, V3, V1);
  }
  VN = dot3(V1, V1);
  if (VN > ORC_EPSDIV0 * ORC_EPSDIV0) {
    double l = sqrt(VN);
    V1[0] /= l; V1[1] /= l; V1[2] /= l;
  } else
    return 3;
  cross(V3, V1, V2);
  return 0;
}

/* strainAndStressUtils.f90:437-481; T column-major 2x2 = [CA, -SA, SA, CA] */
int orc_shell_stress_trans(const double VX[3], const double VZ[3], double T[4])
{
  double V1[3], V2[3], CA, SA, s;
  globalized_x(VZ, V1);
  if (dot3(V1, V1) <= ORC_EPSDIV0 * ORC_EPSDIV0) return -1;
  cross(VX, V1, V2);
  CA = dot3(V1, VX);
  s = sqrt(dot3(V2, V2));
  SA = dot3(V2, VZ) >= 0.0 ? fabs(s) : -fabs(s); /* Fortran sign(a,b) */
  T[0] = CA;  /* T(1,1) */
  T[2] = SA;  /* T(1,2) */
  T[1] = -SA; /* T(2,1) */
  T[3] = CA;  /* T(2,2) */
  return 0;
}

/* strainAndStressUtils.f90:244-294 */
static void quad4_shape_der(double xi, double eta, const double xL[4], const double yL[4],
                            double sx[4], double sy[4])
{
  double sxi[4], seta[4], j11, j12, j21, j22, det, i11, i12, i21, i22;
  sxi[0] = -(1.0 - eta) * 0.25;
  sxi[1] = (1.0 - eta) * 0.25;
  sxi[2] = (1.0 + eta) * 0.25;
  sxi[3] = -(1.0 + eta) * 0.25;
  seta[0] = -(1.0 - xi) * 0.25;
  seta[1] = -(1.0 + xi) * 0.25;
  seta[2] = (1.0 + xi) * 0.25;
  seta[3] = (1.0 - xi) * 0.25;
  j11 = j12 = j21 = j22 = 0.0;
  for (int k = 0; k < 4; k++) {
    j11 += sxi[k] * xL[k];
    j12 += sxi[k] * yL[k];
    j21 += seta[k] * xL[k];
    j22 += seta[k] * yL[k];
  }
  det = j11 * j22 - j21 * j12;
  i11 = j22 / det;
  i22 = j11 / det;
  i12 = -j12 / det;
  i21 = -j21 / det;
  for (int k = 0; k < 4; k++) {
    sx[k] = i11 * sxi[k] + i12 * seta[k];
    sy[k] = i21 * sxi[k] + i22 * seta[k];
  }
}

/* strainAndStressUtils.f90:165-217.  T_el column-major 3x3 (row i = axis i);
 * B_el(3,nndof,4) column-major: B_el[(c) + 3*(d) + 3*nndof*(n)] */
void orc_strain_disp_quad4(int nndof, const double *xEl, const double *yEl, const double *zEl,
                           const double T_el[9], double xi, double eta, double zPos,
                           double *B_el)
{
  double xL[4], yL[4], sx[4], sy[4], vec[3];
#define BE(c, d, n) B_el[((c)-1) + 3 * ((d)-1) + 3 * nndof * ((n)-1)]
#define TEL(i, j) T_el[((i)-1) + 3 * ((j)-1)]
  memset(B_el, 0, sizeof(double) * 3 * nndof * 4);
  xL[0] = 0.0;
  yL[0] = 0.0;
  for (int in = 1; in < 4; in++) {
    vec[0] = xEl[in] - xEl[0];
    vec[1] = yEl[in] - yEl[0];
    vec[2] = zEl[in] - zEl[0];
    xL[in] = TEL(1, 1) * vec[0] + TEL(1, 2) * vec[1] + TEL(1, 3) * vec[2];
    yL[in] = TEL(2, 1) * vec[0] + TEL(2, 2) * vec[1] + TEL(2, 3) * vec[2];
  }
  quad4_shape_der(xi, eta, xL, yL, sx, sy);
  for (int n = 1; n <= 4; n++) {
    BE(1, 1, n) = sx[n - 1];
    BE(3, 1, n) = sy[n - 1];
    BE(2, 2, n) = sy[n - 1];
    BE(3, 2, n) = sx[n - 1];
  }
  if (nndof >= 5 && fabs(zPos) > ORC_EPSDIV0)
    for (int n = 1; n <= 4; n++)
      for (int c = 1; c <= 3; c++) {
        BE(c, 4, n) = -zPos * BE(c, 2, n);
        BE(c, 5, n) = zPos * BE(c, 1, n);
      }
#undef BE
}

/* strainAndStressUtils.f90:101-162 */
void orc_strain_disp_cst(int nndof, const double *xEl, const double *yEl, const double *zEl,
                         const double T_el[9], double zPos, double *B_el)
{
  double xLij[3][3], yLij[3][3], vec[3], a2;
#define BE(c, d, n) B_el[((c)-1) + 3 * ((d)-1) + 3 * nndof * ((n)-1)]
  memset(B_el, 0, sizeof(double) * 3 * nndof * 3);
  memset(xLij, 0, sizeof(xLij));
  memset(yLij, 0, sizeof(yLij));
  for (int in = 0; in < 3; in++)
    for (int jn = 0; jn < 3; jn++) {
      if (in == jn) continue;
      vec[0] = xEl[in] - xEl[jn];
      vec[1] = yEl[in] - yEl[jn];
      vec[2] = zEl[in] - zEl[jn];
      xLij[in][jn] = TEL(1, 1) * vec[0] + TEL(1, 2) * vec[1] + TEL(1, 3) * vec[2];
      yLij[in][jn] = TEL(2, 1) * vec[0] + TEL(2, 2) * vec[1] + TEL(2, 3) * vec[2];
    }
  a2 = xLij[1][0] * yLij[2][0] - xLij[2][0] * yLij[1][0];
  BE(1, 1, 1) = yLij[1][2] / a2;
  BE(1, 1, 2) = yLij[2][0] / a2;
  BE(1, 1, 3) = yLij[0][1] / a2;
  BE(3, 1, 1) = -xLij[1][2] / a2;
  BE(3, 1, 2) = -xLij[2][0] / a2;
  BE(3, 1, 3) = -xLij[0][1] / a2;
  BE(2, 2, 1) = -xLij[1][2] / a2;
  BE(2, 2, 2) = -xLij[2][0] / a2;
  BE(2, 2, 3) = -xLij[0][1] / a2;
  BE(3, 2, 1) = yLij[1][2] / a2;
  BE(3, 2, 2) = yLij[2][0] / a2;
  BE(3, 2, 3) = yLij[0][1] / a2;
  if (nndof >= 5 && fabs(zPos) > ORC_EPSDIV0)
    for (int n = 1; n <= 3; n++)
      for (int c = 1; c <= 3; c++) {
        BE(c, 4, n) = -zPos * BE(c, 2, n);
        BE(c, 5, n) = zPos * BE(c, 1, n);
      }
#undef BE
#undef TEL
}

/* elStressModule.f90:841-847 */
static void tra_strain(double eps[3], const double T_str[4])
{
  eps[2] = 0.5 * eps[2];
  orc_rotate2d(eps, T_str, eps);
  eps[2] = 2.0 * eps[2];
}

/* -ffqStressForm of the legacy FFQ shell (type 22): 2 = 2x2 Gauss points (default, what STR24 always uses), 1 = 1x1 */
static int g_ffq_stress_form = 2;
void orc_set_ffq_stress_form(int form) { g_ffq_stress_form = form; }
int orc_get_ffq_stress_form(void) { return g_ffq_stress_form; }

/* elStressModule.f90:738-849 (STR22a).  Outputs column-major: SR(6,4), SS(6,4), sigma(3,8), epsil(3,8). */
static void str22a(int nGauss, const double *XG, const double *YG, const double *ZG, const double *THK,
                   const double Cmat[9], const double T_el[9], const double T_str[4],
                   const double *EV, double *SR, double *SS, double *sigma, double *epsil)
{
  enum { nenod = 4, nedof = 24 };
  static const int iClose[4] = {1, 2, 2, 1}, iFar[4] = {2, 1, 1, 2};
  static const int jClose[4] = {1, 1, 2, 2}, jFar[4] = {2, 2, 1, 1};
  const double sqrt3 = sqrt(3.0);
  const double f1 = 0.5 + 0.5 * sqrt3, f2 = 0.5 - 0.5 * sqrt3;
  double gauss[2], hHalf, vld[nedof], B_L[3 * nedof], B_U[3 * nedof];
  double epsGU[2][2][3], epsGL[2][2][3], epsU[3], epsL[3], sigU[3], sigL[3];
  int i, j, n;

  if (nGauss == 1) {
    gauss[0] = 0.0;
    gauss[1] = 1.0;
  } else {
    gauss[0] = -1.0 / sqrt3;
    gauss[1] = 1.0 / sqrt3;
  }

  hHalf = (THK[0] + THK[1] + THK[2] + THK[3]) / (double)(2 * nenod);
  for (i = 1; i <= nGauss; i++)
    for (j = 1; j <= nGauss; j++) {
      orc_strain_disp_quad4(6, XG, YG, ZG, T_el, gauss[i - 1], gauss[j - 1], hHalf, B_U);
      orc_strain_disp_quad4(6, XG, YG, ZG, T_el, gauss[i - 1], gauss[j - 1], -hHalf, B_L);
      if (i == 1 && j == 1)
        for (int i1 = 0; i1 < 8; i1++)
          for (int r = 0; r < 3; r++)
            vld[3 * i1 + r] = T_el[r + 0] * EV[3 * i1] + T_el[r + 3] * EV[3 * i1 + 1] +
                              T_el[r + 6] * EV[3 * i1 + 2];
      for (int c = 0; c < 3; c++) {
        double su = 0.0, sl = 0.0;
        for (int k = 0; k < nedof; k++) {
          su += B_U[c + 3 * k] * vld[k];
          sl += B_L[c + 3 * k] * vld[k];
        }
        epsGU[i - 1][j - 1][c] = su;
        epsGL[i - 1][j - 1][c] = sl;
      }
      tra_strain(epsGU[i - 1][j - 1], T_str);
      tra_strain(epsGL[i - 1][j - 1], T_str);
    }
  /* after the loops the Fortran DO variable i equals nGauss+1 (used as thk(i) below) */
  i = nGauss + 1;

  for (n = 1; n <= nenod; n++) {
    int i1 = iClose[n - 1], j1 = jClose[n - 1], i2 = iFar[n - 1], j2 = jFar[n - 1];
    for (int c = 0; c < 3; c++) {
      if (nGauss == 1) {
        epsU[c] = epsGU[0][0][c];
        epsL[c] = epsGL[0][0][c];
      } else {
        epsU[c] = f1 * epsGU[i1 - 1][j1 - 1][c] + f2 * epsGU[i2 - 1][j2 - 1][c];
        epsL[c] = f1 * epsGL[i1 - 1][j1 - 1][c] + f2 * epsGL[i2 - 1][j2 - 1][c];
      }
    }
    for (int c = 0; c < 3; c++) {
      epsil[c + 3 * (n - 1)] = epsU[c];
      epsil[c + 3 * (nenod + n - 1)] = epsL[c];
    }
    for (int c = 0; c < 3; c++) {
      sigU[c] = Cmat[c] * epsU[0] + Cmat[c + 3] * epsU[1] + Cmat[c + 6] * epsU[2];
      sigL[c] = Cmat[c] * epsL[0] + Cmat[c + 3] * epsL[1] + Cmat[c + 6] * epsL[2];
    }
    for (int c = 0; c < 3; c++) {
      double sigMem = (sigU[c] + sigL[c]) * 0.5, sigBen = (sigU[c] - sigL[c]) * 0.5;
      sigma[c + 3 * (n - 1)] = sigU[c];
      sigma[c + 3 * (nenod + n - 1)] = sigL[c];
      SR[c + 6 * (n - 1)] = sigMem * THK[i - 1];
      SR[3 + c + 6 * (n - 1)] = sigBen * THK[i - 1] * THK[i - 1] / 6.0;
    }
    for (int c = 0; c < 6; c++) SS[c + 6 * (n - 1)] = 0.0;
  }
}

/* elStressModule.f90:1005-1078 (EF branch not taken: fedem_stress passes no nodal forces
 * unless -nodalForces).  EV is overwritten by its projection like the reference. */
static int str24_worker(int nGauss, const double xg[4], const double yg[4], const double zg[4], double emod,
                        double rny, const double thk[4], double ev[24], double SR[24], double SS[24],
                        double sigma[24], double epsil[24])
{
  double Cmat[9], T_el[9], T_str[4], PMAT[24 * 24], tmp[24], V1[3], V2[3], V3[3];
  int ierr;

  orc_iso_mat2d(emod, rny, Cmat);
  if (orc_pmat_stiff(4, xg, yg, zg, PMAT) < 0) return 1;

  for (int i = 0; i < 24; i++) {
    double s = 0.0;
    for (int j = 0; j < 24; j++) s += PMAT[i + 24 * j] * ev[j];
    tmp[i] = s;
  }
  memcpy(ev, tmp, sizeof(tmp));

  ierr = orc_shell_element_axes(4, xg, yg, zg, V1, V2, V3);
  if (ierr != 0) return ierr;
  for (int j = 0; j < 3; j++) { /* T_el(1,:)=V1, T_el(2,:)=V2, T_el(3,:)=V3 */
    T_el[0 + 3 * j] = V1[j];
    T_el[1 + 3 * j] = V2[j];
    T_el[2 + 3 * j] = V3[j];
  }
  ierr = orc_shell_stress_trans(V1, V3, T_str);
  if (ierr != 0) return ierr;

  str22a(nGauss, xg, yg, zg, thk, Cmat, T_el, T_str, ev, SR, SS, sigma, epsil);
  return 0;
}

int orc_str24(const double xg[4], const double yg[4], const double zg[4], double emod,
              double rny, const double thk[4], double ev[24], double SR[24], double SS[24],
              double sigma[24], double epsil[24])
{
  return str24_worker(2, xg, yg, zg, emod, rny, thk, ev, SR, SS, sigma, epsil);
}

/* STR22 (elStressModule.f90:640-733) for -ffqStressForm 1 or 2: the same projection (lStiffProj = .true.) and STR22a with 1x1 or
 * 2x2 Gauss points; -ffqStressForm 0 (STR22b, Femlib nodal evaluation) is not restated: returns -99. */
int orc_str22(const double xg[4], const double yg[4], const double zg[4], double emod, double rny, const double thk[4],
              double ev[24], double SR[24], double SS[24], double sigma[24], double epsil[24])
{
  if (g_ffq_stress_form != 1 && g_ffq_stress_form != 2) return -99;
  return str24_worker(g_ffq_stress_form, xg, yg, zg, emod, rny, thk, ev, SR, SS, sigma, epsil);
}
