/* solids_lin.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): the linear solid elements of fedem_stress.
 *   STR44  8-node hexahedron   src/vpmStress/elStressModule.f90:1590-1689 -> HEXA32 / JABN30 / LINHEX
 *                              (src/Femlib/hexa.f:278-413,802-880,1116-1156); compatible element only
 *                              (-useIncompatibleModes is a private option that defaults to off)
 *   STR45  4-node tetrahedron  elStressModule.f90:1695-1727 -> CSTetStrain / cstetbmat / pdvcoor / cstetvolume
 *                              (src/Femlib/cstetra.f90:23-116,385-452,528-615)
 *   STR46  6-node wedge        elStressModule.f90:1733-1764 -> Ipri6Strain / ipri6bmat / pdvn / ipri6extrapolH
 *                              (src/Femlib/ipri6.f90:392-533,614-816) + JACI31 (src/Femlib/jaci31.f)
 * Component order of the reference: HEX8 (xx,yy,zz,xy,xz,yz) like the quadratic solids, TET4 and WEDG6
 * (xx,yy,zz,xy,yz,zx) -- their B-matrix rows 5 and 6 are swapped with respect to the others and the stress
 * module stores them as they come.  1-based index macros so that the code reads like the Fortran.
 * No reference golden exists for these routines: parity unpinned, patch-tested in tests/test_oracle_cpu.py. */
#include <math.h>
#include <string.h>
#include "oracle.h"

#define A2(a, i, j, ld) a[((i)-1) + (ld) * ((j)-1)]

static void iso_mat3d(double Emod, double Rnu, double *C) /* isoMatModule.f90:63-91 */
{
  double fac = Emod / ((1.0 + Rnu) * (1.0 - Rnu - Rnu));
  memset(C, 0, 36 * sizeof(double));
  A2(C, 1, 1, 6) = (1.0 - Rnu) * fac;
  A2(C, 2, 1, 6) = Rnu * fac;
  A2(C, 3, 1, 6) = A2(C, 2, 1, 6);
  A2(C, 1, 2, 6) = A2(C, 2, 1, 6); A2(C, 2, 2, 6) = A2(C, 1, 1, 6); A2(C, 3, 2, 6) = A2(C, 2, 1, 6);
  A2(C, 1, 3, 6) = A2(C, 2, 1, 6); A2(C, 2, 3, 6) = A2(C, 2, 1, 6); A2(C, 3, 3, 6) = A2(C, 1, 1, 6);
  A2(C, 4, 4, 6) = (0.5 - Rnu) * fac;
  A2(C, 5, 5, 6) = A2(C, 4, 4, 6);
  A2(C, 6, 6, 6) = A2(C, 4, 4, 6);
}

static void iso_mat3d_inv(double Emod, double Rnu, double *C) /* isoMatModule.f90:95-120 */
{
  memset(C, 0, 36 * sizeof(double));
  A2(C, 1, 1, 6) = 1.0 / Emod;
  A2(C, 2, 1, 6) = -Rnu / Emod;
  A2(C, 3, 1, 6) = A2(C, 2, 1, 6);
  A2(C, 1, 2, 6) = A2(C, 2, 1, 6); A2(C, 2, 2, 6) = A2(C, 1, 1, 6); A2(C, 3, 2, 6) = A2(C, 2, 1, 6);
  A2(C, 1, 3, 6) = A2(C, 2, 1, 6); A2(C, 2, 3, 6) = A2(C, 2, 1, 6); A2(C, 3, 3, 6) = A2(C, 1, 1, 6);
  A2(C, 4, 4, 6) = 2.0 * (1.0 + Rnu) / Emod;
  A2(C, 5, 5, 6) = A2(C, 4, 4, 6);
  A2(C, 6, 6, 6) = A2(C, 4, 4, 6);
}

static void matvec6(const double *C, const double *x, double *y)
{
  for (int i = 1; i <= 6; i++) {
    double s = 0.0;
    for (int k = 1; k <= 6; k++) s += A2(C, i, k, 6) * x[k - 1];
    y[i - 1] = s;
  }
}

/* ---- TET4 ------------------------------------------------------------------------------------ */
int orc_str45(const double *x, const double *y, const double *z, double emod, double rny, const double *v,
              double *sigma /* (6,4) */, double *epsil /* (6,4) */)
{
#define X(i) x[(i)-1]
#define Y(i) y[(i)-1]
#define Z(i) z[(i)-1]
  double E[36], a[4], b[4], c[4], bmat[6 * 12], s12[3], s13[3], s14[3], cr[3], volume, factor;
  iso_mat3d(emod, rny, E);
  s12[0] = X(2) - X(1); s12[1] = Y(2) - Y(1); s12[2] = Z(2) - Z(1);
  s13[0] = X(3) - X(1); s13[1] = Y(3) - Y(1); s13[2] = Z(3) - Z(1);
  s14[0] = X(4) - X(1); s14[1] = Y(4) - Y(1); s14[2] = Z(4) - Z(1);
  cr[0] = s12[1] * s13[2] - s12[2] * s13[1];
  cr[1] = s12[2] * s13[0] - s12[0] * s13[2];
  cr[2] = s12[0] * s13[1] - s12[1] * s13[0];
  volume = (cr[0] * s14[0] + cr[1] * s14[1] + cr[2] * s14[2]) / 6.0;
  if (volume > ORC_EPSDIV0) factor = 1.0 / (6.0 * volume);
  else return 1; /* cstetbmat ierr = -1 -> STR45 ierr = 1 */
  /* pdvcoor */
  a[0] = (Y(2) - Y(3)) * (Z(4) - Z(2)) - (Y(4) - Y(2)) * (Z(2) - Z(3));
  a[1] = (Y(4) - Y(3)) * (Z(1) - Z(3)) - (Y(1) - Y(3)) * (Z(4) - Z(3));
  a[2] = (Y(4) - Y(1)) * (Z(2) - Z(4)) - (Y(2) - Y(4)) * (Z(4) - Z(1));
  a[3] = (Y(2) - Y(1)) * (Z(3) - Z(1)) - (Y(3) - Y(1)) * (Z(2) - Z(1));
  b[0] = (Z(2) - Z(3)) * (X(4) - X(2)) - (Z(4) - Z(2)) * (X(2) - X(3));
  b[1] = (Z(4) - Z(3)) * (X(1) - X(3)) - (Z(1) - Z(3)) * (X(4) - X(3));
  b[2] = (Z(4) - Z(1)) * (X(2) - X(4)) - (Z(2) - Z(4)) * (X(4) - X(1));
  b[3] = (Z(2) - Z(1)) * (X(3) - X(1)) - (Z(3) - Z(1)) * (X(2) - X(1));
  c[0] = (X(2) - X(3)) * (Y(4) - Y(2)) - (X(4) - X(2)) * (Y(2) - Y(3));
  c[1] = (X(4) - X(3)) * (Y(1) - Y(3)) - (X(1) - X(3)) * (Y(4) - Y(3));
  c[2] = (X(4) - X(1)) * (Y(2) - Y(4)) - (X(2) - X(4)) * (Y(4) - Y(1));
  c[3] = (X(2) - X(1)) * (Y(3) - Y(1)) - (X(3) - X(1)) * (Y(2) - Y(1));
  memset(bmat, 0, sizeof(bmat));
  for (int inod = 1; inod <= 4; inod++) {
    int xpos = (inod - 1) * 3 + 1, ypos = xpos + 1, zpos = xpos + 2;
    double ai = a[inod - 1] * factor, bi = b[inod - 1] * factor, ci = c[inod - 1] * factor;
    A2(bmat, 1, xpos, 6) = ai; A2(bmat, 2, ypos, 6) = bi; A2(bmat, 3, zpos, 6) = ci;
    A2(bmat, 4, xpos, 6) = bi; A2(bmat, 4, ypos, 6) = ai;
    A2(bmat, 5, ypos, 6) = ci; A2(bmat, 5, zpos, 6) = bi;
    A2(bmat, 6, zpos, 6) = ai; A2(bmat, 6, xpos, 6) = ci;
  }
  for (int i = 1; i <= 6; i++) {
    double s = 0.0;
    for (int k = 1; k <= 12; k++) s += A2(bmat, i, k, 6) * v[k - 1];
    epsil[i - 1] = s;
  }
  matvec6(E, epsil, sigma);
  for (int n = 2; n <= 4; n++) {
    memcpy(sigma + 6 * (n - 1), sigma, 6 * sizeof(double));
    memcpy(epsil + 6 * (n - 1), epsil, 6 * sizeof(double));
  }
  return 0;
#undef X
#undef Y
#undef Z
}

/* JACI31 for MEK nodes; returns 0 or -1 */
static int jaci31(double *ji, double *detj, const double *dnxi, const double *dnet, const double *dnze, const double *xg,
                  const double *yg, const double *zg, int mek)
{
  double J[9];
  const double eps = DBL_MIN * 100.0;
  memset(J, 0, sizeof(J));
  for (int i = 1; i <= mek; i++) {
    A2(J, 1, 1, 3) += dnxi[i - 1] * xg[i - 1]; A2(J, 1, 2, 3) += dnxi[i - 1] * yg[i - 1]; A2(J, 1, 3, 3) += dnxi[i - 1] * zg[i - 1];
    A2(J, 2, 1, 3) += dnet[i - 1] * xg[i - 1]; A2(J, 2, 2, 3) += dnet[i - 1] * yg[i - 1]; A2(J, 2, 3, 3) += dnet[i - 1] * zg[i - 1];
    A2(J, 3, 1, 3) += dnze[i - 1] * xg[i - 1]; A2(J, 3, 2, 3) += dnze[i - 1] * yg[i - 1]; A2(J, 3, 3, 3) += dnze[i - 1] * zg[i - 1];
  }
#define JJ(i, j) A2(J, i, j, 3)
#define JI(i, j) A2(ji, i, j, 3)
  *detj = JJ(1, 1) * (JJ(2, 2) * JJ(3, 3) - JJ(2, 3) * JJ(3, 2)) + JJ(1, 2) * (JJ(2, 3) * JJ(3, 1) - JJ(2, 1) * JJ(3, 3)) +
          JJ(1, 3) * (JJ(2, 1) * JJ(3, 2) - JJ(2, 2) * JJ(3, 1));
  if (fabs(*detj) - eps <= 0.0) return -1;
  JI(1, 1) = (JJ(2, 2) * JJ(3, 3) - JJ(2, 3) * JJ(3, 2)) / *detj;
  JI(1, 2) = (JJ(3, 2) * JJ(1, 3) - JJ(3, 3) * JJ(1, 2)) / *detj;
  JI(1, 3) = (JJ(1, 2) * JJ(2, 3) - JJ(1, 3) * JJ(2, 2)) / *detj;
  JI(2, 1) = (JJ(3, 1) * JJ(2, 3) - JJ(3, 3) * JJ(2, 1)) / *detj;
  JI(2, 2) = (JJ(1, 1) * JJ(3, 3) - JJ(1, 3) * JJ(3, 1)) / *detj;
  JI(2, 3) = (JJ(2, 1) * JJ(1, 3) - JJ(2, 3) * JJ(1, 1)) / *detj;
  JI(3, 1) = (JJ(2, 1) * JJ(3, 2) - JJ(2, 2) * JJ(3, 1)) / *detj;
  JI(3, 2) = (JJ(3, 1) * JJ(1, 2) - JJ(3, 2) * JJ(1, 1)) / *detj;
  JI(3, 3) = (JJ(1, 1) * JJ(2, 2) - JJ(1, 2) * JJ(2, 1)) / *detj;
  *detj = fabs(*detj);
  return 0;
#undef JJ
#undef JI
}

/* ---- WEDG6 ----------------------------------------------------------------------------------- */
static int ipri6bmat(const double *x, const double *y, const double *z, const double *xi, double zeta, double *bmat)
{
  double a[6], b[6], c[6], inja[9], detjac;
  memset(bmat, 0, 6 * 18 * sizeof(double));
  /* pdvn */
  a[0] = (1.0 - zeta) * 0.5; a[3] = (1.0 + zeta) * 0.5; a[2] = -a[0]; a[5] = -a[3]; a[1] = 0.0; a[4] = 0.0;
  b[1] = a[0]; b[4] = a[3]; b[2] = -a[0]; b[5] = -a[3]; b[0] = 0.0; b[3] = 0.0;
  for (int i = 0; i < 3; i++) { c[i] = -xi[i] * 0.5; c[3 + i] = xi[i] * 0.5; }
  if (jaci31(inja, &detjac, a, b, c, x, y, z, 6) < 0) return -1;
  for (int inod = 1; inod <= 6; inod++) {
    int xpos = (inod - 1) * 3 + 1, ypos = xpos + 1, zpos = xpos + 2;
    A2(bmat, 1, xpos, 6) = a[inod - 1] * A2(inja, 1, 1, 3) + b[inod - 1] * A2(inja, 1, 2, 3) + c[inod - 1] * A2(inja, 1, 3, 3);
    A2(bmat, 2, ypos, 6) = a[inod - 1] * A2(inja, 2, 1, 3) + b[inod - 1] * A2(inja, 2, 2, 3) + c[inod - 1] * A2(inja, 2, 3, 3);
    A2(bmat, 3, zpos, 6) = a[inod - 1] * A2(inja, 3, 1, 3) + b[inod - 1] * A2(inja, 3, 2, 3) + c[inod - 1] * A2(inja, 3, 3, 3);
    A2(bmat, 4, xpos, 6) = A2(bmat, 2, ypos, 6);
    A2(bmat, 4, ypos, 6) = A2(bmat, 1, xpos, 6);
    A2(bmat, 5, ypos, 6) = A2(bmat, 3, zpos, 6);
    A2(bmat, 5, zpos, 6) = A2(bmat, 2, ypos, 6);
    A2(bmat, 6, xpos, 6) = A2(bmat, 3, zpos, 6);
    A2(bmat, 6, zpos, 6) = A2(bmat, 1, xpos, 6);
  }
  return 0;
}

int orc_str46(const double *x, const double *y, const double *z, double emod, double rny, int stressForm, const double *v,
              double *sigma /* (6,6) */, double *epsil /* (6,6) */)
{
  double E[36], bmat[6 * 18], aux[36], spntxi[9], spntzeta[2];
  int code = stressForm == 1 ? 3 : stressForm;
  iso_mat3d(emod, rny, E);
#define SX(i, j) A2(spntxi, i, j, 3)
#define ST(c, p) A2(epsil, c, p, 6)
#define AX(c, p) A2(aux, c, p, 6)
  switch (code) {
  case 0:
    spntzeta[0] = -1.0; spntzeta[1] = 1.0;
    for (int k = 0; k < 9; k++) spntxi[k] = 0.0;
    for (int i = 1; i <= 3; i++) SX(i, i) = 1.0;
    break;
  case 1:
    spntzeta[0] = -1.0 / sqrt(3.0); spntzeta[1] = -spntzeta[0];
    for (int k = 0; k < 9; k++) spntxi[k] = 1.0 / 6.0;
    for (int i = 1; i <= 3; i++) SX(i, i) = 2.0 / 3.0;
    break;
  case 2:
    spntzeta[0] = -1.0 / sqrt(3.0); spntzeta[1] = -spntzeta[0];
    for (int k = 0; k < 9; k++) spntxi[k] = 0.5;
    for (int i = 1; i <= 3; i++) SX(i, i) = 0.0;
    break;
  case 3:
    spntzeta[0] = spntzeta[1] = 0.0;
    for (int k = 0; k < 9; k++) spntxi[k] = 0.5;
    for (int i = 1; i <= 3; i++) SX(i, i) = 0.0;
    break;
  default:
    return 1;
  }
  for (int zpnt = 1; zpnt <= 2; zpnt++) {
    for (int xpnt = 1; xpnt <= 3; xpnt++) {
      if (ipri6bmat(x, y, z, &SX(1, xpnt), spntzeta[zpnt - 1], bmat) < 0) return 1;
      for (int i = 1; i <= 6; i++) {
        double s = 0.0;
        for (int k = 1; k <= 18; k++) s += A2(bmat, i, k, 6) * v[k - 1];
        ST(i, xpnt + 3 * zpnt - 3) = s;
      }
    }
    if (code == 3) break;
  }
  if (code == 1 || code == 2) {
    for (int c = 1; c <= 6; c++) {
      if (code == 1) {
        AX(c, 1) = (5.0 * ST(c, 1) - ST(c, 2) - ST(c, 3)) / 3.0;
        AX(c, 2) = (5.0 * ST(c, 2) - ST(c, 1) - ST(c, 3)) / 3.0;
        AX(c, 3) = (5.0 * ST(c, 3) - ST(c, 1) - ST(c, 2)) / 3.0;
        AX(c, 4) = (5.0 * ST(c, 4) - ST(c, 5) - ST(c, 6)) / 3.0;
        AX(c, 5) = (5.0 * ST(c, 5) - ST(c, 4) - ST(c, 6)) / 3.0;
        AX(c, 6) = (5.0 * ST(c, 6) - ST(c, 4) - ST(c, 5)) / 3.0;
      } else {
        AX(c, 1) = ST(c, 2) + ST(c, 3) - ST(c, 1);
        AX(c, 2) = ST(c, 1) + ST(c, 3) - ST(c, 2);
        AX(c, 3) = ST(c, 1) + ST(c, 2) - ST(c, 3);
        AX(c, 4) = ST(c, 5) + ST(c, 6) - ST(c, 4);
        AX(c, 5) = ST(c, 4) + ST(c, 6) - ST(c, 5);
        AX(c, 6) = ST(c, 4) + ST(c, 5) - ST(c, 6);
      }
    }
    { /* ipri6extrapolH */
      double zm1 = 0.5 * (sqrt(3.0) - 1.0), zp1 = 0.5 * (sqrt(3.0) + 1.0);
      for (int c = 1; c <= 6; c++)
        for (int p = 1; p <= 3; p++) {
          ST(c, p) = zp1 * AX(c, p) - zm1 * AX(c, p + 3);
          ST(c, p + 3) = zp1 * AX(c, p + 3) - zm1 * AX(c, p);
        }
    }
  } else if (code == 3) {
    for (int c = 1; c <= 6; c++) {
      ST(c, 4) = ST(c, 2) + ST(c, 3) - ST(c, 1);
      ST(c, 5) = ST(c, 1) + ST(c, 3) - ST(c, 2);
      ST(c, 6) = ST(c, 1) + ST(c, 2) - ST(c, 3);
      for (int p = 1; p <= 3; p++) ST(c, p) = ST(c, p + 3);
    }
  }
  for (int p = 1; p <= 6; p++) matvec6(E, &ST(1, p), &A2(sigma, 1, p, 6));
  return 0;
#undef SX
#undef ST
#undef AX
}

/* ---- HEX8 ------------------------------------------------------------------------------------ */
static const double CXI[8] = {-1., 1., 1., -1., -1., 1., 1., -1.};
static const double CETA[8] = {-1., -1., 1., 1., -1., -1., 1., 1.};
static const double CZETA[8] = {-1., -1., -1., -1., 1., 1., 1., 1.};

/* JABN30 with IOP = 1: inverse Jacobian BJ and DETJ; DETJ <= 0 leaves BJ untouched */
static void jabn30(double *BJ, const double *X, const double *Y, const double *Z, double XI, double ETA, double ZETA, double *DETJ)
{
#define XX(i) X[(i)-1]
#define YY(i) Y[(i)-1]
#define ZZ(i) Z[(i)-1]
  double A1 = 1. + XI, A2_ = 1. - XI, B1 = 1. + ETA, B2 = 1. - ETA, C1 = 1. + ZETA, C2 = 1. - ZETA;
  double J11, J12, J13, J21, J22, J23, J31, J32, J33, DET;
  memset(BJ, 0, 9 * sizeof(double));
  J11 = 0.125 * (B1 * C1 * (XX(7) - XX(8)) - B2 * C1 * (XX(5) - XX(6)) + B1 * C2 * (XX(3) - XX(4)) - B2 * C2 * (XX(1) - XX(2)));
  J12 = 0.125 * (B1 * C1 * (YY(7) - YY(8)) - B2 * C1 * (YY(5) - YY(6)) + B1 * C2 * (YY(3) - YY(4)) - B2 * C2 * (YY(1) - YY(2)));
  J13 = 0.125 * (B1 * C1 * (ZZ(7) - ZZ(8)) - B2 * C1 * (ZZ(5) - ZZ(6)) + B1 * C2 * (ZZ(3) - ZZ(4)) - B2 * C2 * (ZZ(1) - ZZ(2)));
  J21 = 0.125 * (A1 * C1 * (XX(7) - XX(6)) + A2_ * C1 * (XX(8) - XX(5)) + A1 * C2 * (XX(3) - XX(2)) + A2_ * C2 * (XX(4) - XX(1)));
  J22 = 0.125 * (A1 * C1 * (YY(7) - YY(6)) + A2_ * C1 * (YY(8) - YY(5)) + A1 * C2 * (YY(3) - YY(2)) + A2_ * C2 * (YY(4) - YY(1)));
  J23 = 0.125 * (A1 * C1 * (ZZ(7) - ZZ(6)) + A2_ * C1 * (ZZ(8) - ZZ(5)) + A1 * C2 * (ZZ(3) - ZZ(2)) + A2_ * C2 * (ZZ(4) - ZZ(1)));
  J31 = 0.125 * (A1 * B1 * (XX(7) - XX(3)) + A2_ * B1 * (XX(8) - XX(4)) + A2_ * B2 * (XX(5) - XX(1)) + A1 * B2 * (XX(6) - XX(2)));
  J32 = 0.125 * (A1 * B1 * (YY(7) - YY(3)) + A2_ * B1 * (YY(8) - YY(4)) + A2_ * B2 * (YY(5) - YY(1)) + A1 * B2 * (YY(6) - YY(2)));
  J33 = 0.125 * (A1 * B1 * (ZZ(7) - ZZ(3)) + A2_ * B1 * (ZZ(8) - ZZ(4)) + A2_ * B2 * (ZZ(5) - ZZ(1)) + A1 * B2 * (ZZ(6) - ZZ(2)));
  *DETJ = J11 * (J22 * J33 - J23 * J32) - J12 * (J21 * J33 - J23 * J31) + J13 * (J21 * J32 - J22 * J31);
  if (*DETJ <= 0.0) return;
  DET = 1. / *DETJ;
  A2(BJ, 1, 1, 3) = DET * (J22 * J33 - J32 * J23);
  A2(BJ, 1, 2, 3) = -DET * (J12 * J33 - J13 * J32);
  A2(BJ, 1, 3, 3) = DET * (J12 * J23 - J13 * J22);
  A2(BJ, 2, 1, 3) = -DET * (J21 * J33 - J23 * J31);
  A2(BJ, 2, 2, 3) = DET * (J11 * J33 - J13 * J31);
  A2(BJ, 2, 3, 3) = -DET * (J11 * J23 - J13 * J21);
  A2(BJ, 3, 1, 3) = DET * (J21 * J32 - J22 * J31);
  A2(BJ, 3, 2, 3) = -DET * (J11 * J32 - J12 * J31);
  A2(BJ, 3, 3, 3) = DET * (J11 * J22 - J12 * J21);
#undef XX
#undef YY
#undef ZZ
}

/* HEXA32 with IOP = 0: SI(6,24) = E . P at the stress point ABC*(XI,ETA,ZETA); returns IER */
static int hexa32(double *SI, double *DETJ, const double *E, const double *X, const double *Y, const double *Z, double XI, double ETA,
                  double ZETA, int LOP)
{
  double BJ[9], P[6 * 24], ABC, XXI, XETA, XZETA;
  if (LOP == 2) ABC = sqrt(1.0 / 3.0);
  else if (LOP == 3) ABC = sqrt(0.6);
  else ABC = 1.0;
  XXI = ABC * XI; XETA = ABC * ETA; XZETA = ABC * ZETA;
  memset(P, 0, sizeof(P));
  memset(SI, 0, 6 * 24 * sizeof(double));
  jabn30(BJ, X, Y, Z, XXI, XETA, XZETA, DETJ);
  if (*DETJ < 0.0) return 1;
  for (int I = 1; I <= 8; I++) {
    int L = 3 * I - 2;
    for (int J = 1; J <= 3; J++) {
      int K = 3 * I + J - 3;
      A2(P, J, K, 6) = 0.125 * (A2(BJ, J, 1, 3) * CXI[I - 1] * (1. + XETA * CETA[I - 1]) * (1. + XZETA * CZETA[I - 1]) +
                                A2(BJ, J, 2, 3) * CETA[I - 1] * (1. + XXI * CXI[I - 1]) * (1. + XZETA * CZETA[I - 1]) +
                                A2(BJ, J, 3, 3) * CZETA[I - 1] * (1. + XXI * CXI[I - 1]) * (1. + XETA * CETA[I - 1]));
    }
    A2(P, 4, L, 6) = A2(P, 2, L + 1, 6);
    A2(P, 4, L + 1, 6) = A2(P, 1, L, 6);
    A2(P, 5, L, 6) = A2(P, 3, L + 2, 6);
    A2(P, 5, L + 2, 6) = A2(P, 1, L, 6);
    A2(P, 6, L + 1, 6) = A2(P, 3, L + 2, 6);
    A2(P, 6, L + 2, 6) = A2(P, 2, L + 1, 6);
  }
  for (int I = 1; I <= 6; I++)
    for (int J = 1; J <= 24; J++)
      for (int K = 1; K <= 6; K++) A2(SI, I, J, 6) += A2(E, I, K, 6) * A2(P, K, J, 6);
  return 0;
}

int orc_str44(const double *x, const double *y, const double *z, double emod, double rny, int stressForm, const double *v,
              double *sigma /* (6,8) */, double *epsil /* (6,8) */)
{
  double E[36], Einv[36], SI[6 * 24], si[6 * 8], sig[6] = {0, 0, 0, 0, 0, 0}, vol = 0.0, detJ;
  const int LOP = stressForm == 0 ? 0 : 2; /* NIP - 1 */
  const double sqrt3 = sqrt(3.0);
  iso_mat3d(emod, rny, E);
  iso_mat3d_inv(emod, rny, Einv);
  for (int n = 1; n <= 8; n++) {
    if (hexa32(SI, &detJ, E, x, y, z, CXI[n - 1], CETA[n - 1], CZETA[n - 1], LOP) != 0) return 1;
    for (int i = 1; i <= 6; i++) {
      double s = 0.0;
      for (int k = 1; k <= 24; k++) s += A2(SI, i, k, 6) * v[k - 1];
      A2(sigma, i, n, 6) = s;
    }
    if (stressForm == 0) matvec6(Einv, &A2(sigma, 1, n, 6), &A2(epsil, 1, n, 6));
    else if (stressForm == 1) {
      for (int i = 0; i < 6; i++) sig[i] += A2(sigma, i + 1, n, 6) * detJ;
      vol += detJ;
    }
  }
  if (stressForm == 1) {
    for (int i = 0; i < 6; i++) sigma[i] = sig[i] / vol;
    matvec6(Einv, sigma, epsil);
    for (int n = 2; n <= 8; n++) {
      memcpy(&A2(sigma, 1, n, 6), sigma, 6 * sizeof(double));
      memcpy(&A2(epsil, 1, n, 6), epsil, 6 * sizeof(double));
    }
  } else if (stressForm >= 2) {
    /* LINHEX node table (hexa.f:1137-1139) */
    static const double LX[8] = {1., -1., -1., 1., 1., -1., -1., 1.};
    static const double LY[8] = {1., 1., -1., -1., 1., 1., -1., -1.};
    static const double LZ[8] = {1., 1., 1., 1., -1., -1., -1., -1.};
    memset(si, 0, sizeof(si));
    for (int n = 1; n <= 8; n++) {
      double xx = -CXI[n - 1] * sqrt3, yy = -CETA[n - 1] * sqrt3, zz = -CZETA[n - 1] * sqrt3;
      for (int i = 1; i <= 8; i++) {
        double w = 0.125 * (1.0 + xx * LX[i - 1]) * (1.0 + yy * LY[i - 1]) * (1.0 + zz * LZ[i - 1]);
        for (int c = 1; c <= 6; c++) A2(si, c, n, 6) += A2(sigma, c, i, 6) * w;
      }
    }
    memcpy(sigma, si, sizeof(si));
    for (int n = 1; n <= 8; n++) matvec6(Einv, &A2(sigma, 1, n, 6), &A2(epsil, 1, n, 6));
  }
  return 0;
}

/* ---- WEDG15 ---------------------------------------------------------------------------------------
 * STR42 (src/vpmStress/elStressModule.f90:1381-1465) -> IPRI32 (src/Femlib/ipri.f:391-767) -> DN1531 (ipri.f:2867-2971),
 * JACI31.  Component order (xx,yy,zz,xy,xz,yz).  The Fortran 77 source writes its abscissae as default-REAL literals
 * (ZE = +-.577350269189626 in an IMPLICIT DOUBLE PRECISION unit is still a REAL*4 constant): reproduced with float casts. */
static void dn1531(double *DNL1, double *DNL2, double *DNZE, double RL1, double RL2, double RL3, double ZE)
{
  double RL1RL1 = RL1 * RL1, RL1RL2 = RL1 * RL2, RL1RL3 = RL1 * RL3, RL1ZE = RL1 * ZE, RL2RL2 = RL2 * RL2, RL2RL3 = RL2 * RL3,
         RL2ZE = RL2 * ZE, RL3RL3 = RL3 * RL3, RL3ZE = RL3 * ZE, ZEZE = ZE * ZE;
#define D1_(i) DNL1[(i)-1]
#define D2_(i) DNL2[(i)-1]
#define DZ_(i) DNZE[(i)-1]
  D1_(1) = -1. + 0.5 * ZE + 0.5 * ZEZE + 2. * RL1 - 2. * RL1ZE;
  D1_(2) = 2. * RL2 - 2. * RL2ZE;
  D1_(3) = 0.;
  D1_(4) = -2. * RL2 + 2. * RL2ZE;
  D1_(5) = 1. - 0.5 * ZE - 0.5 * ZEZE - 2. * RL3 + 2. * RL3ZE;
  D1_(6) = 2. * (1. - ZE) * (RL3 - RL1);
  D1_(7) = 1. - ZEZE;
  D1_(8) = 0.;
  D1_(9) = -1. + ZEZE;
  D1_(10) = -1. - 0.5 * ZE + 0.5 * ZEZE + 2. * RL1 + 2. * RL1ZE;
  D1_(11) = 2. * RL2 + 2. * RL2ZE;
  D1_(12) = 0.;
  D1_(13) = -2. * RL2 - 2. * RL2ZE;
  D1_(14) = 1. + 0.5 * ZE - 0.5 * ZEZE - 2. * RL3 - 2. * RL3ZE;
  D1_(15) = 2. * (1. + ZE) * (RL3 - RL1);
  D2_(1) = 0.;
  D2_(2) = 2. * RL1 - 2. * RL1ZE;
  D2_(3) = -1. + 0.5 * ZE + 0.5 * ZEZE + 2. * RL2 - 2. * RL2ZE;
  D2_(4) = 2. * (RL3 - RL2) * (1. - ZE);
  D2_(5) = 1. - 0.5 * ZE - 0.5 * ZEZE - 2. * RL3 + 2. * RL3ZE;
  D2_(6) = -2. * RL1 + 2. * RL1ZE;
  D2_(7) = 0.;
  D2_(8) = 1. - ZEZE;
  D2_(9) = -1. + ZEZE;
  D2_(10) = 0.;
  D2_(11) = 2. * RL1 + 2. * RL1ZE;
  D2_(12) = -1. - 0.5 * ZE + 0.5 * ZEZE + 2. * RL2 + 2. * RL2ZE;
  D2_(13) = 2. * (RL3 - RL2) * (1. + ZE);
  D2_(14) = 1. + 0.5 * ZE - 0.5 * ZEZE - 2. * RL3 - 2. * RL3ZE;
  D2_(15) = -2. * RL1 - 2. * RL1ZE;
  DZ_(1) = 0.5 * RL1 + RL1ZE - RL1RL1;
  DZ_(2) = -2. * RL1RL2;
  DZ_(3) = 0.5 * RL2 + RL2ZE - RL2RL2;
  DZ_(4) = -2. * RL2RL3;
  DZ_(5) = 0.5 * RL3 + RL3ZE - RL3RL3;
  DZ_(6) = -2. * RL1RL3;
  DZ_(7) = -2. * RL1ZE;
  DZ_(8) = -2. * RL2ZE;
  DZ_(9) = -2. * RL3ZE;
  DZ_(10) = -0.5 * RL1 + RL1ZE + RL1RL1;
  DZ_(11) = 2. * RL1RL2;
  DZ_(12) = -0.5 * RL2 + RL2ZE + RL2RL2;
  DZ_(13) = 2. * RL2RL3;
  DZ_(14) = -0.5 * RL3 + RL3ZE + RL3RL3;
  DZ_(15) = 2. * RL1RL3;
#undef D1_
#undef D2_
#undef DZ_
}

/* IPRI32 for NSTRP in {3, 6}, NSTRPZ in {2, 3}, IOPXP = IOPE0 = 0: SIG(6, NSTRP*NSTRPZ) */
static int ipri32(double *SIG, const double *V, const double *XG, const double *YG, const double *ZG, double YOUNG, double RNY,
                  int NSTRP, int NSTRPZ)
{
  double RL1[7] = {0}, RL2[7] = {0}, RL3[7] = {0}, ZE[3] = {0}, DNL1[15], DNL2[15], DNZE[15], JI[9], DETJ, B[3], DB[15][6][3];
  double D = YOUNG * (1. - RNY) / ((1. + RNY) * (1. - 2. * RNY));
  double D1 = D * RNY / (1. - RNY);
  double D2 = D * (1. - 2. * RNY) / (2. * (1. - RNY));
  int N = 0;
  memset(SIG, 0, 6 * 21 * sizeof(double));
  if (NSTRP == 3) {
    RL1[0] = .5; RL2[0] = .5; RL3[0] = .0;
    RL1[1] = .0; RL2[1] = .5; RL3[1] = .5;
    RL1[2] = .5; RL2[2] = .0; RL3[2] = .5;
  } else if (NSTRP == 6) {
    RL1[0] = 1.0; RL1[1] = 0.5; RL2[1] = 0.5; RL2[2] = 1.0; RL2[3] = 0.5; RL3[3] = 0.5; RL3[4] = 1.0; RL1[5] = 0.5; RL3[5] = 0.5;
  } else
    return -1;
  if (NSTRPZ == 2) { ZE[0] = (double)-.577350269189626f; ZE[1] = (double).577350269189626f; }
  else if (NSTRPZ == 3) {
    ZE[0] = (double)-.774596669241483f; ZE[2] = (double).774596669241483f;
    if (NSTRP == 6) { ZE[0] = -1.0; ZE[2] = 1.0; }
  } else
    return -1;
  for (int L = 1; L <= NSTRPZ; L++)
    for (int K = 1; K <= NSTRP; K++) {
      dn1531(DNL1, DNL2, DNZE, RL1[K - 1], RL2[K - 1], RL3[K - 1], ZE[L - 1]);
      N = N + 1;
      if (jaci31(JI, &DETJ, DNL1, DNL2, DNZE, XG, YG, ZG, 15) < 0) return -1;
      for (int J = 1; J <= 15; J++) {
        B[0] = A2(JI, 1, 1, 3) * DNL1[J - 1] + A2(JI, 1, 2, 3) * DNL2[J - 1] + A2(JI, 1, 3, 3) * DNZE[J - 1];
        B[1] = A2(JI, 2, 1, 3) * DNL1[J - 1] + A2(JI, 2, 2, 3) * DNL2[J - 1] + A2(JI, 2, 3, 3) * DNZE[J - 1];
        B[2] = A2(JI, 3, 1, 3) * DNL1[J - 1] + A2(JI, 3, 2, 3) * DNL2[J - 1] + A2(JI, 3, 3, 3) * DNZE[J - 1];
        DB[J - 1][0][0] = D * B[0];  DB[J - 1][1][0] = D1 * B[0]; DB[J - 1][2][0] = DB[J - 1][1][0];
        DB[J - 1][3][0] = D2 * B[1]; DB[J - 1][4][0] = D2 * B[2]; DB[J - 1][5][0] = 0.0;
        DB[J - 1][0][1] = D1 * B[1]; DB[J - 1][1][1] = D * B[1];  DB[J - 1][2][1] = DB[J - 1][0][1];
        DB[J - 1][3][1] = D2 * B[0]; DB[J - 1][4][1] = 0.0;       DB[J - 1][5][1] = DB[J - 1][4][0];
        DB[J - 1][0][2] = D1 * B[2]; DB[J - 1][1][2] = DB[J - 1][0][2]; DB[J - 1][2][2] = D * B[2];
        DB[J - 1][3][2] = 0.0;       DB[J - 1][4][2] = DB[J - 1][3][1]; DB[J - 1][5][2] = DB[J - 1][3][0];
      }
      for (int I = 1; I <= 15; I++) {
        int I1 = (I - 1) * 3 + 1, I2 = I1 + 1, I3 = I2 + 1;
        for (int J = 1; J <= 6; J++)
          A2(SIG, J, N, 6) = A2(SIG, J, N, 6) + DB[I - 1][J - 1][0] * V[I1 - 1] + DB[I - 1][J - 1][1] * V[I2 - 1] + DB[I - 1][J - 1][2] * V[I3 - 1];
      }
    }
  return 0;
}

int orc_str42(const double *x, const double *y, const double *z, double emod, double rny, int stressForm, const double *v,
              double *sigma /* (6,15) */, double *epsil /* (6,15) */)
{
  double Einv[36], SIGG[6 * 21];
  const double zm1 = 0.5 * sqrt(3.0) - 0.5, zp1 = zm1 + 1.0;
  int n = stressForm == 0 ? 6 : 3, nZ = stressForm == 0 ? 3 : 2;
  iso_mat3d_inv(emod, rny, Einv);
  if (ipri32(SIGG, v, x, y, z, emod, rny, n, nZ) != 0) return 1;
#define SG(c, p) A2(SIGG, c, p, 6)
#define SI(c, p) A2(sigma, c, p, 6)
#define EP(c, p) A2(epsil, c, p, 6)
  for (int c = 1; c <= 6; c++) {
    if (stressForm == 0) {
      for (int p = 1; p <= 7; p++) SI(c, p) = SG(c, p);
      SI(c, 8) = SG(c, 9);
      SI(c, 9) = SG(c, 11);
      for (int p = 10; p <= 15; p++) SI(c, p) = SG(c, p + 3);
    } else {
      EP(c, 1) = SG(c, 3) + SG(c, 1) - SG(c, 2);
      EP(c, 2) = SG(c, 2) + SG(c, 1) - SG(c, 3);
      EP(c, 3) = SG(c, 2) + SG(c, 3) - SG(c, 1);
      EP(c, 4) = SG(c, 6) + SG(c, 4) - SG(c, 5);
      EP(c, 5) = SG(c, 5) + SG(c, 4) - SG(c, 6);
      EP(c, 6) = SG(c, 5) + SG(c, 6) - SG(c, 4);
      SI(c, 1) = zp1 * EP(c, 1) - zm1 * EP(c, 4);
      SI(c, 3) = zp1 * EP(c, 2) - zm1 * EP(c, 5);
      SI(c, 5) = zp1 * EP(c, 3) - zm1 * EP(c, 6);
      SI(c, 10) = zp1 * EP(c, 4) - zm1 * EP(c, 1);
      SI(c, 12) = zp1 * EP(c, 5) - zm1 * EP(c, 2);
      SI(c, 14) = zp1 * EP(c, 6) - zm1 * EP(c, 3);
      SI(c, 2) = 0.5 * (SI(c, 1) + SI(c, 3));
      SI(c, 4) = 0.5 * (SI(c, 3) + SI(c, 5));
      SI(c, 6) = 0.5 * (SI(c, 5) + SI(c, 1));
      SI(c, 7) = 0.5 * (SI(c, 1) + SI(c, 10));
      SI(c, 8) = 0.5 * (SI(c, 3) + SI(c, 12));
      SI(c, 9) = 0.5 * (SI(c, 5) + SI(c, 14));
      SI(c, 11) = 0.5 * (SI(c, 10) + SI(c, 12));
      SI(c, 13) = 0.5 * (SI(c, 12) + SI(c, 14));
      SI(c, 15) = 0.5 * (SI(c, 14) + SI(c, 10));
    }
  }
  for (int p = 1; p <= 15; p++) matvec6(Einv, &SI(1, p), &EP(1, p));
  return 0;
#undef SG
#undef SI
#undef EP
}
