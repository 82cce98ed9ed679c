/* thickshell.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): the subparametric thick shells,
 * 6-noded triangle (type 31) and 8-noded quadrilateral (type 32), 5 local DOFs per node.
 *
 * Follows, statement by statement:
 *   STR31 / STR32      src/vpmStress/elStressModule.f90:1083-1180, 1186-1302
 *   SCTS30 / SCTS32    src/Femlib/scts.f:1241-1292, 1806-2052   (stress matrix of the triangle)
 *   SCQS30 / SCQS32    src/Femlib/scqs.f:770-824, 1345-1604     (stress matrix of the quadrilateral)
 *   DNI630, JACO30, LNCS30, RCOS30, CHQT30   src/Femlib/scts.f:53-110, 346-455, 460-560, 1086-1235, 7-52
 *   DNI830, CHQA30     src/Femlib/scqs.f:55-118, 7-54
 *   isoMat2Dinv        src/Femlib/isoMatModule.f90:42-59
 *   tratensor -> FFaTensorTransforms::rotate3D  fedem-foundation/src/FFaLib/FFaAlgebra/FFaTensorTransforms.C:369-405
 *
 * The Femlib sources are fixed-form Fortran 77 with default-REAL literals: "1.2", ".333333333", "1.E-06", "1.E-10",
 * "1.E-03" are REAL*4 constants promoted to double, reproduced here with float casts (1.2f is NOT 1.2).
 *
 * This is the one element family for which the reference holds a unit-level known-answer test
 * (src/vpmStress/vpmStressTests/testThickShell.pf, geometry in vpmStressTests/ffl.f90); tests/test_thickshell_cpu.py
 * re-expresses its eight cases against orc_el_stress-level output with the reference's tolerance (1e-15).
 *
 * LAMBI(node, r, c): r = 1,2 the two in-surface axes of the node system, r = 3 the normal; c = global component. */
#include "oracle.h"
#include <math.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------------------------- */
/* LNCS30 (scts.f:460-560): local x', y' from the normal RET(:,3) as prescribed by ICODIR.  RET is 3x3 column-major. */
#define RET(i, j) ret[((i) - 1) + 3 * ((j) - 1)]
static void lncs30(double *ret, int icodir, int *ierr)
{
  double da, rl1, rl2;
  int stage = icodir;
  if (icodir < 1 || icodir > 3) { *ierr = -1; return; }
  if (stage == 1) {
    /* x'-z' plane parallel to the z'-X plane */
    da = sqrt(RET(2, 3) * RET(2, 3) + RET(3, 3) * RET(3, 3));
    if (da - (double)1.E-03f < 0.0) {
      *ierr = 1;
      stage = 2;
    } else {
      RET(1, 2) = 0.;
      RET(2, 2) = RET(3, 3);
      RET(3, 2) = -RET(2, 3);
      RET(1, 1) = RET(3, 3) * RET(3, 3) + RET(2, 3) * RET(2, 3);
      RET(2, 1) = -RET(2, 3) * RET(1, 3);
      RET(3, 1) = -RET(1, 3) * RET(3, 3);
      goto normalize;
    }
  }
  if (stage == 2) {
    /* x'-z' plane parallel to the z'-Y plane */
    da = sqrt(RET(1, 3) * RET(1, 3) + RET(3, 3) * RET(3, 3));
    if (da - (double)1.E-03f < 0.0) {
      *ierr = 2;
      stage = 3;
    } else {
      RET(1, 2) = -RET(3, 3);
      RET(2, 2) = 0.;
      RET(3, 2) = RET(1, 3);
      RET(1, 1) = -RET(1, 3) * RET(2, 3);
      RET(2, 1) = RET(3, 3) * RET(3, 3) + RET(1, 3) * RET(1, 3);
      RET(3, 1) = -RET(3, 3) * RET(2, 3);
      goto normalize;
    }
  }
  /* x'-z' plane parallel to the z'-Z plane */
  da = sqrt(RET(1, 3) * RET(1, 3) + RET(2, 3) * RET(2, 3));
  if (da - (double)1.E-03f < 0.0) { *ierr = -3; return; }
  RET(1, 2) = RET(2, 3);
  RET(2, 2) = -RET(1, 3);
  RET(3, 2) = 0.;
  RET(1, 1) = -RET(1, 3) * RET(3, 3);
  RET(2, 1) = -RET(2, 3) * RET(3, 3);
  RET(3, 1) = RET(2, 3) * RET(2, 3) + RET(1, 3) * RET(1, 3);
normalize:
  rl1 = sqrt(RET(1, 1) * RET(1, 1) + RET(2, 1) * RET(2, 1) + RET(3, 1) * RET(3, 1));
  rl2 = sqrt(RET(1, 2) * RET(1, 2) + RET(2, 2) * RET(2, 2) + RET(3, 2) * RET(3, 2));
  for (int i = 1; i <= 3; i++) {
    RET(i, 1) = RET(i, 1) / rl1;
    RET(i, 2) = RET(i, 2) / rl2;
  }
}
#undef RET

/* RCOS30 (scts.f:1086-1235): direction cosines of the node system of node n (0-based). lambi[node][r][c]. */
static void rcos30(const double *xg, const double *yg, const double *zg, const double *dnl1, const double *dnl2,
                   double (*lambi)[3][3], int n, int mek, int icodir, int *ierr)
{
  double dxdl1 = 0., dydl1 = 0., dzdl1 = 0., dxdl2 = 0., dydl2 = 0., dzdl2 = 0., r, r1, r2, tet[9], v3[3];
  for (int k = 0; k < mek; k++) {
    dxdl1 = dxdl1 + xg[k] * dnl1[k];
    dydl1 = dydl1 + yg[k] * dnl1[k];
    dzdl1 = dzdl1 + zg[k] * dnl1[k];
    dxdl2 = dxdl2 + xg[k] * dnl2[k];
    dydl2 = dydl2 + yg[k] * dnl2[k];
    dzdl2 = dzdl2 + zg[k] * dnl2[k];
  }
  v3[0] = dydl1 * dzdl2 - dydl2 * dzdl1;
  v3[1] = dxdl2 * dzdl1 - dxdl1 * dzdl2;
  v3[2] = dxdl1 * dydl2 - dxdl2 * dydl1;
  r = sqrt(v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2]);
  if (r - 1.0e-15 <= 0.0) { *ierr = -1; return; }
  r = 1 / r;
  if (icodir > 0) {
    memset(tet, 0, sizeof(tet));
    tet[6] = v3[0] * r;
    tet[7] = v3[1] * r;
    tet[8] = v3[2] * r;
    lncs30(tet, icodir, ierr);
    if (*ierr < 0) return;
    for (int c = 0; c < 3; c++) {
      lambi[n][0][c] = tet[c];       /* LAMBI(N,1,c) = TET(c,1) */
      lambi[n][1][c] = tet[3 + c];   /* LAMBI(N,2,c) = TET(c,2) */
      lambi[n][2][c] = tet[6 + c];   /* LAMBI(N,3,c) = TET(c,3) */
    }
  } else {
    lambi[n][2][0] = v3[0] * r;
    lambi[n][2][1] = v3[1] * r;
    lambi[n][2][2] = v3[2] * r;
    r1 = sqrt(dxdl1 * dxdl1 + dydl1 * dydl1 + dzdl1 * dzdl1);
    r1 = 1 / r1;
    lambi[n][0][0] = dxdl1 * r1;
    lambi[n][0][1] = dydl1 * r1;
    lambi[n][0][2] = dzdl1 * r1;
    tet[3] = v3[1] * dzdl1 - dydl1 * v3[2];
    tet[4] = v3[2] * dxdl1 - v3[0] * dzdl1;
    tet[5] = v3[0] * dydl1 - dxdl1 * v3[1];
    r2 = sqrt(tet[3] * tet[3] + tet[4] * tet[4] + tet[5] * tet[5]);
    r2 = 1 / r2;
    lambi[n][1][0] = tet[3] * r2;
    lambi[n][1][1] = tet[4] * r2;
    lambi[n][1][2] = tet[5] * r2;
  }
}

/* JACO30 (scts.f:346-455).  J, JI 3x3 column-major (J(i,j) = j_[(i-1)+3*(j-1)]). */
#define J_(i, j) ja[((i) - 1) + 3 * ((j) - 1)]
#define JI_(i, j) ji[((i) - 1) + 3 * ((j) - 1)]
static void jaco30(double *ja, double *ji, double *detj, const double *dn1, const double *dn2, const double *n,
                   const double *xg, const double *yg, const double *zg, const double *th, double (*lambi)[3][3],
                   double ze, int mek, int *ierr)
{
  double zet, f1, f2, f3, f4, detji;
  for (int k = 0; k < 9; k++) ja[k] = 0.0;
  for (int i = 0; i < mek; i++) {
    zet = ze * th[i] * 0.5;
    f1 = xg[i] + zet * lambi[i][2][0];
    f2 = yg[i] + zet * lambi[i][2][1];
    f3 = zg[i] + zet * lambi[i][2][2];
    f4 = n[i] * .5 * th[i];
    J_(1, 1) = J_(1, 1) + dn1[i] * f1;
    J_(1, 2) = J_(1, 2) + dn1[i] * f2;
    J_(1, 3) = J_(1, 3) + dn1[i] * f3;
    J_(2, 1) = J_(2, 1) + dn2[i] * f1;
    J_(2, 2) = J_(2, 2) + dn2[i] * f2;
    J_(2, 3) = J_(2, 3) + dn2[i] * f3;
    J_(3, 1) = J_(3, 1) + f4 * lambi[i][2][0];
    J_(3, 2) = J_(3, 2) + f4 * lambi[i][2][1];
    J_(3, 3) = J_(3, 3) + f4 * lambi[i][2][2];
  }
  f1 = J_(2, 2) * J_(3, 3) - J_(2, 3) * J_(3, 2);
  f2 = J_(2, 3) * J_(3, 1) - J_(2, 1) * J_(3, 3);
  f3 = J_(2, 1) * J_(3, 2) - J_(2, 2) * J_(3, 1);
  *detj = J_(1, 1) * f1 + J_(1, 2) * f2 + J_(1, 3) * f3;
  if (fabs(*detj) - (double)1.E-10f <= 0.0) { *ierr = -3; return; }
  detji = 1 / *detj;
  *detj = fabs(*detj);
  JI_(1, 1) = detji * f1;
  JI_(2, 1) = detji * f2;
  JI_(3, 1) = detji * f3;
  JI_(1, 2) = detji * (J_(3, 2) * J_(1, 3) - J_(3, 3) * J_(1, 2));
  JI_(1, 3) = detji * (J_(1, 2) * J_(2, 3) - J_(1, 3) * J_(2, 2));
  JI_(2, 2) = detji * (J_(1, 1) * J_(3, 3) - J_(1, 3) * J_(3, 1));
  JI_(2, 3) = detji * (J_(2, 1) * J_(1, 3) - J_(2, 3) * J_(1, 1));
  JI_(3, 2) = detji * (J_(3, 1) * J_(1, 2) - J_(3, 2) * J_(1, 1));
  JI_(3, 3) = detji * (J_(1, 1) * J_(2, 2) - J_(1, 2) * J_(2, 1));
}

/* DNI630 (scts.f:53-110) */
static void dni630(double *dnl1, double *dnl2, double *nl12, double rl1, double rl2, int lin)
{
  if (lin - 1 > 0) {
    nl12[0] = 2.0 * rl1 * rl1 - rl1;
    nl12[1] = 2.0 * rl2 * rl2 - rl2;
    nl12[2] = 2.0 * rl1 * rl1 + 2.0 * rl2 * rl2 + 4.0 * rl1 * rl2 - 3.0 * rl1 - 3.0 * rl2 + 1.0;
    nl12[3] = 4.0 * rl1 * rl2;
    nl12[4] = 4.0 * rl2 - 4.0 * rl1 * rl2 - 4.0 * rl2 * rl2;
    nl12[5] = 4.0 * rl1 - 4.0 * rl1 * rl1 - 4.0 * rl1 * rl2;
  } else if (lin - 1 == 0) {
    /* least-square linearised shape functions */
    nl12[0] = 0.6 * rl1 - 0.2;
    nl12[1] = 0.6 * rl2 - 0.2;
    nl12[2] = -0.6 * rl1 - 0.6 * rl2 + 0.4;
    nl12[3] = 0.8 * rl1 + 0.8 * rl2 - 0.2;
    nl12[4] = -0.8 * rl1 + 0.6;
    nl12[5] = -0.8 * rl2 + 0.6;
  }
  dnl1[0] = 4.0 * rl1 - 1.0;
  dnl2[0] = 0.0;
  dnl1[1] = 0.0;
  dnl2[1] = 4.0 * rl2 - 1.0;
  dnl1[2] = 4.0 * rl1 + 4.0 * rl2 - 3.0;
  dnl2[2] = dnl1[2];
  dnl1[3] = 4.0 * rl2;
  dnl2[3] = 4.0 * rl1;
  dnl1[4] = -4.0 * rl2;
  dnl2[4] = 4.0 - 4.0 * rl1 - 8.0 * rl2;
  dnl1[5] = 4.0 - 8.0 * rl1 - 4.0 * rl2;
  dnl2[5] = -4.0 * rl1;
}

/* DNI830 (scqs.f:55-118); the third output is the shape function itself */
static void dni830(double *dnxi, double *dnet, double *dnze, double xi, double et, const double *xii, const double *eti)
{
  for (int i = 0; i < 8; i++) {
    double xixi = xii[i] * xi, eteti = eti[i] * et, ets, xis;
    if ((i & 1) == 0) {          /* corner nodes 1,3,5,7 */
      ets = .25 * (1. + eteti);
      xis = .25 * (1. + xixi);
      dnxi[i] = xii[i] * (2. * xixi + eteti) * ets;
      dnet[i] = eti[i] * xis * (2. * eteti + xixi);
      dnze[i] = xis * (1. + eteti) * (xixi + eteti - 1.);
    } else if (i == 1 || i == 5) { /* mid-side nodes with xii = 0 */
      ets = 1. + eteti;
      xis = .5 * (1. - xi * xi);
      dnxi[i] = -xi * ets;
      dnet[i] = eti[i] * xis;
      dnze[i] = ets * xis;
    } else {                     /* mid-side nodes with eti = 0 */
      ets = (1. - et * et) * .5;
      xis = 1. + xixi;
      dnxi[i] = xii[i] * ets;
      dnet[i] = -xis * et;
      dnze[i] = xis * ets;
    }
  }
}

/* CHQT30 (scts.f:7-52): mid-side nodes must divide their edge in a ratio inside (1:3, 3:1) */
static void chqt30(const double *xg, const double *yg, const double *zg, int *ierr)
{
  int n = 1, n1 = 4;
  for (int i = 1; i <= 3; i++) {
    double dx = xg[n1 - 1] - xg[n - 1], dy = yg[n1 - 1] - yg[n - 1], dz = zg[n1 - 1] - zg[n - 1], rl, rl1, rl2;
    rl1 = sqrt(dx * dx + dy * dy + dz * dz);
    if (n - 2 > 0) n = 0;
    dx = xg[n] - xg[n1 - 1];
    dy = yg[n] - yg[n1 - 1];
    dz = zg[n] - zg[n1 - 1];
    rl2 = sqrt(dx * dx + dy * dy + dz * dz);
    if (rl1 - (double)1.E-06f <= 0.0 || rl2 - (double)1.E-06f <= 0.0) { *ierr = -1; return; }
    rl = rl1 / rl2;
    if (rl - (double).333333333f < 0.0) { *ierr = -1; return; }   /* IF(...)40,30,30 */
    if (rl - 3. >= 0.0) { *ierr = -1; return; }
    n = n + 1;
    n1 = n1 + 1;
  }
}

/* CHQA30 (scqs.f:7-54) */
static void chqa30(const double *xg, const double *yg, const double *zg, int *ierr)
{
  int n = 1;
  for (int i = 1; i <= 4; i++) {
    int n1 = n + 1;
    double dx = xg[n1 - 1] - xg[n - 1], dy = yg[n1 - 1] - yg[n - 1], dz = zg[n1 - 1] - zg[n - 1], rl, rl1, rl2;
    rl1 = sqrt(dx * dx + dy * dy + dz * dz);
    if (n - 6 > 0) n = -1;
    dx = xg[n + 1] - xg[n1 - 1];
    dy = yg[n + 1] - yg[n1 - 1];
    dz = zg[n + 1] - zg[n1 - 1];
    rl2 = sqrt(dx * dx + dy * dy + dz * dz);
    if (rl1 - (double)1.E-06f <= 0.0 || rl2 - (double)1.E-06f <= 0.0) { *ierr = -1; return; }
    rl = rl1 / rl2;
    if (rl - (double).333333333f <= 0.0) { *ierr = -1; return; }  /* IF(...)40,40,30 */
    if (rl - 3. >= 0.0) { *ierr = -1; return; }
    n = n + 2;
  }
}

/* The common body of SCTS32 (scts.f:1838-2015) and SCQS32 (scqs.f:1395-1575) after the shape function derivatives:
 * SIG is (5, 5*mek) column-major, LAMP 3x3 column-major. */
#define LAMP(i, j) lamp[((i) - 1) + 3 * ((j) - 1)]
#define A_(i, j) a[(i) - 1][(j) - 1]
static void stress_matrix(double *sig, double *lamp, const double *xg, const double *yg, const double *zg,
                          const double *th, double young, double rny, const double *dn1, const double *dn2,
                          const double *nn, double ze, double (*lambi)[3][3], int icodir, int mek, int *ierr)
{
  double D11, D12, D33, D44, ja[9], ji[9], detj, rl, a[3][3];
  D11 = young / (1. - rny * rny);
  D12 = D11 * rny;
  D33 = D11 * (1. - rny) * .5;
  D44 = D33 / (double)1.2f;
  jaco30(ja, ji, &detj, dn1, dn2, nn, xg, yg, zg, th, lambi, ze, mek, ierr);
  if (*ierr < 0) return;
  /* surface normal in the stress point */
  LAMP(1, 3) = J_(1, 2) * J_(2, 3) - J_(2, 2) * J_(1, 3);
  LAMP(2, 3) = J_(2, 1) * J_(1, 3) - J_(1, 1) * J_(2, 3);
  LAMP(3, 3) = J_(1, 1) * J_(2, 2) - J_(2, 1) * J_(1, 2);
  rl = sqrt(LAMP(1, 3) * LAMP(1, 3) + LAMP(2, 3) * LAMP(2, 3) + LAMP(3, 3) * LAMP(3, 3));
  LAMP(1, 3) = LAMP(1, 3) / rl;
  LAMP(2, 3) = LAMP(2, 3) / rl;
  LAMP(3, 3) = LAMP(3, 3) / rl;
  if (icodir > 0) {
    lncs30(lamp, icodir, ierr);
    if (*ierr < 0) return;
  } else {
    rl = J_(1, 1) * J_(1, 1) + J_(1, 2) * J_(1, 2) + J_(1, 3) * J_(1, 3);
    rl = 1.0 / sqrt(rl);
    for (int k = 1; k <= 3; k++) LAMP(k, 1) = J_(1, k) * rl;
    LAMP(1, 2) = LAMP(2, 3) * J_(1, 3) - J_(1, 2) * LAMP(3, 3);
    LAMP(2, 2) = J_(1, 1) * LAMP(3, 3) - LAMP(1, 3) * J_(1, 3);
    LAMP(3, 2) = LAMP(1, 3) * J_(1, 2) - J_(1, 1) * LAMP(2, 3);
    rl = LAMP(1, 2) * LAMP(1, 2) + LAMP(2, 2) * LAMP(2, 2) + LAMP(3, 2) * LAMP(3, 2);
    rl = 1.0 / sqrt(rl);
    for (int k = 1; k <= 3; k++) LAMP(k, 2) = LAMP(k, 2) * rl;
  }
  A_(1, 1) = LAMP(1, 1) * JI_(1, 1) + LAMP(2, 1) * JI_(2, 1) + LAMP(3, 1) * JI_(3, 1);
  A_(1, 2) = LAMP(1, 1) * JI_(1, 2) + LAMP(2, 1) * JI_(2, 2) + LAMP(3, 1) * JI_(3, 2);
  A_(2, 1) = LAMP(1, 2) * JI_(1, 1) + LAMP(2, 2) * JI_(2, 1) + LAMP(3, 2) * JI_(3, 1);
  A_(2, 2) = LAMP(1, 2) * JI_(1, 2) + LAMP(2, 2) * JI_(2, 2) + LAMP(3, 2) * JI_(3, 2);
  A_(3, 1) = LAMP(1, 3) * JI_(1, 1) + LAMP(2, 3) * JI_(2, 1) + LAMP(3, 3) * JI_(3, 1);
  A_(3, 2) = LAMP(1, 3) * JI_(1, 2) + LAMP(2, 3) * JI_(2, 2) + LAMP(3, 3) * JI_(3, 2);
  A_(3, 3) = LAMP(1, 3) * JI_(1, 3) + LAMP(2, 3) * JI_(2, 3) + LAMP(3, 3) * JI_(3, 3);
  for (int i = 0; i < mek; i++) {
    double B[3], C, A1[5][3], A2[5][2], A3[5][3], A4[5][2], A5[5][3], A6[5][2], A7[5][2];
    B[0] = A_(1, 1) * dn1[i] + A_(1, 2) * dn2[i];
    B[1] = A_(2, 1) * dn1[i] + A_(2, 2) * dn2[i];
    B[2] = A_(3, 1) * dn1[i] + A_(3, 2) * dn2[i];
    C = A_(3, 3) * nn[i];
    memset(A3, 0, sizeof(A3)); memset(A4, 0, sizeof(A4)); memset(A6, 0, sizeof(A6));
    for (int c = 1; c <= 3; c++) {   /* (A1) = (Bi)*(LAMP)' */
      A1[0][c - 1] = LAMP(c, 1) * B[0];
      A1[1][c - 1] = LAMP(c, 2) * B[1];
      A1[2][c - 1] = LAMP(c, 1) * B[1] + LAMP(c, 2) * B[0];
      A1[3][c - 1] = LAMP(c, 3) * B[0] + LAMP(c, 1) * B[2];
      A1[4][c - 1] = LAMP(c, 3) * B[1] + LAMP(c, 2) * B[2];
    }
    for (int ii = 0; ii < 5; ii++) {  /* (A2) = (A1)*(FI) */
      A2[ii][1] = lambi[i][0][0] * A1[ii][0] + lambi[i][0][1] * A1[ii][1] + lambi[i][0][2] * A1[ii][2];
      A2[ii][0] = -lambi[i][1][0] * A1[ii][0] - lambi[i][1][1] * A1[ii][1] - lambi[i][1][2] * A1[ii][2];
    }
    for (int c = 1; c <= 3; c++) {   /* (A3) = (Ci)*(LAMP)' */
      A3[3][c - 1] = LAMP(c, 1) * C;
      A3[4][c - 1] = LAMP(c, 2) * C;
    }
    for (int ii = 3; ii < 5; ii++) {  /* (A4) = (A3)*(FI) */
      A4[ii][1] = lambi[i][0][0] * A3[ii][0] + lambi[i][0][1] * A3[ii][1] + lambi[i][0][2] * A3[ii][2];
      A4[ii][0] = -lambi[i][1][0] * A3[ii][0] - lambi[i][1][1] * A3[ii][1] - lambi[i][1][2] * A3[ii][2];
    }
    for (int c = 0; c < 3; c++) {    /* (A5) = (D)*(A1) */
      A5[0][c] = A1[0][c] * D11 + A1[1][c] * D12;
      A5[1][c] = A1[0][c] * D12 + A1[1][c] * D11;
      A5[2][c] = A1[2][c] * D33;
      A5[3][c] = A1[3][c] * D44;
      A5[4][c] = A1[4][c] * D44;
    }
    for (int c = 0; c < 2; c++) {    /* (A6) = (D)*(A4), (A7) = (D)*(A2) */
      A6[3][c] = A4[3][c] * D44;
      A6[4][c] = A4[4][c] * D44;
      A7[0][c] = A2[0][c] * D11 + A2[1][c] * D12;
      A7[1][c] = A2[0][c] * D12 + A2[1][c] * D11;
      A7[2][c] = A2[2][c] * D33;
      A7[3][c] = A2[3][c] * D44;
      A7[4][c] = A2[4][c] * D44;
    }
    {
      const double thh = .5 * th[i];
      const int nl = i * 5;
      for (int i1 = 0; i1 < 5; i1++) {
        for (int i2 = 0; i2 < 3; i2++) sig[i1 + 5 * (nl + i2)] = A5[i1][i2];
        for (int i3 = 0; i3 < 2; i3++) {
          double v = thh * ze * A7[i1][i3];
          if (i1 >= 3) v = v + thh * A6[i1][i3];
          sig[i1 + 5 * (nl + 3 + i3)] = v;
        }
      }
    }
  }
}
#undef A_
#undef J_
#undef JI_

/* FFaTensorTransforms::rotate3D (FFaTensorTransforms.C:369-405), in place; rotMx = LAMP column-major */
static void rotate3d(double *S, const double *rotMx)
{
  const double *eX = rotMx, *eY = rotMx + 3, *eZ = rotMx + 6;
  double TS11 = eX[0] * S[0] + eY[0] * S[3] + eZ[0] * S[4];
  double TS12 = eX[0] * S[3] + eY[0] * S[1] + eZ[0] * S[5];
  double TS13 = eX[0] * S[4] + eY[0] * S[5] + eZ[0] * S[2];
  double TS21 = eX[1] * S[0] + eY[1] * S[3] + eZ[1] * S[4];
  double TS22 = eX[1] * S[3] + eY[1] * S[1] + eZ[1] * S[5];
  double TS23 = eX[1] * S[4] + eY[1] * S[5] + eZ[1] * S[2];
  double TS31 = eX[2] * S[0] + eY[2] * S[3] + eZ[2] * S[4];
  double TS32 = eX[2] * S[3] + eY[2] * S[1] + eZ[2] * S[5];
  double TS33 = eX[2] * S[4] + eY[2] * S[5] + eZ[2] * S[2];
  S[0] = TS11 * eX[0] + TS12 * eY[0] + TS13 * eZ[0];
  S[1] = TS21 * eX[1] + TS22 * eY[1] + TS23 * eZ[1];
  S[2] = TS31 * eX[2] + TS32 * eY[2] + TS33 * eZ[2];
  S[3] = TS11 * eX[1] + TS12 * eY[1] + TS13 * eZ[1];
  S[4] = TS11 * eX[2] + TS12 * eY[2] + TS13 * eZ[2];
  S[5] = TS21 * eX[2] + TS22 * eY[2] + TS23 * eZ[2];
}
void orc_rotate3d(const double *S, const double *rotMx, double *out)
{
  double t[6];
  memcpy(t, S, sizeof(t));
  rotate3d(t, rotMx);
  memcpy(out, t, sizeof(t));
}

/* local stresses of one sampling point -> global 6-component stress and strain (the common tail of the STR31 / STR32
 * sampling loops, elStressModule.f90:1147-1160, 1259-1272): matmul(SIG,EV), matmul(Einv,.), re-pack, tratensor */
static void point_stress(const double *sig, int nedof, const double *ev5, const double *einv, const double *lamp,
                         double *s6, double *e6)
{
  double s5[5], e5[5];
  for (int r = 0; r < 5; r++) {
    double v = 0.0;
    for (int c = 0; c < nedof; c++) v += sig[r + 5 * c] * ev5[c];
    s5[r] = v;
  }
  for (int r = 0; r < 5; r++) {
    double v = 0.0;
    for (int c = 0; c < 5; c++) v += einv[r + 5 * c] * s5[c];
    e5[r] = v;
  }
  s6[0] = s5[0]; s6[1] = s5[1]; s6[2] = 0.0; s6[3] = s5[2]; s6[4] = s5[3]; s6[5] = s5[4];
  e6[0] = e5[0]; e6[1] = e5[1]; e6[2] = 0.0; e6[3] = e5[2]; e6[4] = e5[3]; e6[5] = e5[4];
  rotate3d(s6, lamp);
  rotate3d(e6, lamp);
}

static void set_einv(double emod, double rny, double *einv)
{
  memset(einv, 0, sizeof(double) * 25);
  einv[0] = 1.0 / emod;
  einv[0 + 5 * 1] = -rny / emod;
  einv[1 + 5 * 0] = einv[0 + 5 * 1];
  einv[1 + 5 * 1] = einv[0];
  einv[2 + 5 * 2] = 2.0 * (1.0 + rny) / emod;
  einv[3 + 5 * 3] = einv[2 + 5 * 2] * 1.2;
  einv[4 + 5 * 4] = einv[3 + 5 * 3];
}

/* local 5-DOF element vector (elStressModule.f90:1128-1131, 1239-1242) */
static void to_local_dofs(int nenod, const double *ev, double (*lambi)[3][3], double *ev5)
{
  for (int i = 0; i < nenod; i++) {
    for (int k = 0; k < 3; k++) ev5[5 * i + k] = ev[6 * i + k];
    for (int r = 0; r < 2; r++)
      ev5[5 * i + 3 + r] = lambi[i][r][0] * ev[6 * i + 3] + lambi[i][r][1] * ev[6 * i + 4] + lambi[i][r][2] * ev[6 * i + 5];
  }
}

/* STR31 (elStressModule.f90:1083-1180).  sigma/epsil (6,12): points 1-6 top surface (3 corners, 3 mid-sides), 7-12 bottom.
 * Returns 0, or 1 when the element is degenerate (warning in the reference: results = hugeVal). */
int orc_str31(const double *xg, const double *yg, const double *zg, double emod, double rny, const double *thk,
              const double *ev, double *sigma, double *epsil)
{
  enum { nenod = 6, nedof = 30 };
  static const double RL1n[6] = {1.0, 0.0, 0.0, 0.5, 0.0, 0.5}, RL2n[6] = {0.0, 1.0, 0.0, 0.5, 0.5, 0.0};
  const double L1[3] = {0.5, 0.0, 0.5}, L2[3] = {0.5, 0.5, 0.0};
  double lambi[6][3][3], lamp[9], sig[5 * 30], einv[25], ev5[30], dnl1[6], dnl2[6], nl12[6];
  int ierr = 0, ip = 0;
  const int icodir = 1, lin = 1;
  set_einv(emod, rny, einv);
  /* SCTS30 */
  for (int i = 0; i < nenod; i++) {
    dni630(dnl1, dnl2, dnl2, RL1n[i], RL2n[i], 0);
    rcos30(xg, yg, zg, dnl1, dnl2, lambi, i, 6, icodir, &ierr);
    if (ierr < 0) return 1;
  }
  to_local_dofs(nenod, ev, lambi, ev5);
  for (int k = 1; k <= 2; k++) {
    const double zeta = (double)(3 - 2 * k);
    for (int i = 0; i < 3; i++) {
      /* SCTS32 */
      ierr = 0;
      chqt30(xg, yg, zg, &ierr);
      if (ierr < 0) return 1;
      dni630(dnl1, dnl2, nl12, L1[i], L2[i], lin);
      stress_matrix(sig, lamp, xg, yg, zg, thk, emod, rny, dnl1, dnl2, nl12, zeta, lambi, icodir, 6, &ierr);
      if (ierr < 0) return 1;
      point_stress(sig, nedof, ev5, einv, lamp, sigma + 6 * (ip + 3 + i), epsil + 6 * (ip + 3 + i));
    }
    /* extrapolate the corner nodes */
    for (int c = 0; c < 6; c++) {
      double *s = sigma + c, *e = epsil + c;
      s[6 * (ip + 0)] = s[6 * (ip + 3)] + s[6 * (ip + 5)] - s[6 * (ip + 4)];
      s[6 * (ip + 1)] = s[6 * (ip + 4)] + s[6 * (ip + 3)] - s[6 * (ip + 5)];
      s[6 * (ip + 2)] = s[6 * (ip + 5)] + s[6 * (ip + 4)] - s[6 * (ip + 3)];
      e[6 * (ip + 0)] = e[6 * (ip + 3)] + e[6 * (ip + 5)] - e[6 * (ip + 4)];
      e[6 * (ip + 1)] = e[6 * (ip + 4)] + e[6 * (ip + 3)] - e[6 * (ip + 5)];
      e[6 * (ip + 2)] = e[6 * (ip + 5)] + e[6 * (ip + 4)] - e[6 * (ip + 3)];
    }
    ip = ip + nenod;
  }
  return 0;
}

/* STR32 (elStressModule.f90:1186-1302).  sigma/epsil (6,16): points 1-8 top surface in node order, 9-16 bottom. */
int orc_str32(const double *xg, const double *yg, const double *zg, double emod, double rny, const double *thk,
              const double *ev, double *sigma, double *epsil)
{
  enum { nenod = 8, nedof = 40 };
  static const double XII[8] = {-1.0, 0.0, 1.0, 1.0, 1.0, 0.0, -1.0, -1.0}, ETI[8] = {-1.0, -1.0, -1.0, 0.0, 1.0, 1.0, 1.0, 0.0};
  const double sqrt3 = sqrt(3.0);
  const double f1 = 0.5 + 0.5 * sqrt3, f2 = 0.5 - 0.5 * sqrt3;
  double lambi[8][3][3], lamp[9], sig[5 * 40], einv[25], ev5[40], dnxi[8], dnet[8], nxiet[8];
  double SigPt[2][2][6], EpsPt[2][2][6];   /* [j][i][c] = SigPt(c,i,j) */
  int ierr = 0, ip = 0;
  const int icodir = 1;
  set_einv(emod, rny, einv);
  /* SCQS30 */
  for (int i = 0; i < nenod; i++) {
    dni830(dnxi, dnet, nxiet, XII[i], ETI[i], XII, ETI);
    rcos30(xg, yg, zg, dnxi, dnet, lambi, i, 8, icodir, &ierr);
    if (ierr < 0) return 1;
  }
  to_local_dofs(nenod, ev, lambi, ev5);
  for (int k = 1; k <= 2; k++) {
    const double zeta = (double)(3 - 2 * k);
    for (int j = 1; j <= 2; j++) {
      const double eta = (double)(2 * j - 3) / sqrt3;
      for (int i = 1; i <= 2; i++) {
        const double xi = (double)(2 * i - 3) / sqrt3;
        /* SCQS32 */
        ierr = 0;
        chqa30(xg, yg, zg, &ierr);
        if (ierr < 0) return 1;
        dni830(dnxi, dnet, nxiet, xi, eta, XII, ETI);
        stress_matrix(sig, lamp, xg, yg, zg, thk, emod, rny, dnxi, dnet, nxiet, zeta, lambi, icodir, 8, &ierr);
        if (ierr < 0) return 1;
        point_stress(sig, nedof, ev5, einv, lamp, SigPt[j - 1][i - 1], EpsPt[j - 1][i - 1]);
      }
    }
    for (int c = 0; c < 6; c++) {
      double *s = sigma + c + 6 * ip, *e = epsil + c + 6 * ip;
      /* extrapolate the corner nodes (1,3,5,7) */
      s[6 * 0] = f1 * SigPt[0][0][c] + f2 * SigPt[1][1][c];
      s[6 * 2] = f1 * SigPt[0][1][c] + f2 * SigPt[1][0][c];
      s[6 * 4] = f1 * SigPt[1][1][c] + f2 * SigPt[0][0][c];
      s[6 * 6] = f1 * SigPt[1][0][c] + f2 * SigPt[0][1][c];
      e[6 * 0] = f1 * EpsPt[0][0][c] + f2 * EpsPt[1][1][c];
      e[6 * 2] = f1 * EpsPt[0][1][c] + f2 * EpsPt[1][0][c];
      e[6 * 4] = f1 * EpsPt[1][1][c] + f2 * EpsPt[0][0][c];
      e[6 * 6] = f1 * EpsPt[1][0][c] + f2 * EpsPt[0][1][c];
      /* interpolate the mid-side nodes */
      for (int i = 2; i <= 8; i += 2) {
        s[6 * (i - 1)] = 0.5 * (s[6 * (i - 2)] + s[6 * (i % 8)]);
        e[6 * (i - 1)] = 0.5 * (e[6 * (i - 2)] + e[6 * (i % 8)]);
      }
    }
    ip = ip + nenod;
  }
  return 0;
}
