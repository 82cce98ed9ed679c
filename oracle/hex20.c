/* hex20.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): 20-node hexahedron STR43.
 * Follows src/vpmStress/elStressModule.f90:1472-1584 (STR43: nodal evaluation for -stressForm 0,
 * else 2x2x2 Gauss points extrapolated tri-linearly), src/Femlib/ihex.f:224-560 (IHEX32),
 * :2433-2545 (DN2031 shape-function derivatives, FEDEM node order), src/Femlib/jaci31.f (JACI31).
 * The Gauss abscissa .577350269189626 (ihex.f:364-365) is a REAL*4 literal promoted to double. */
#include "oracle.h"
#include <math.h>
#include <string.h>

static const double XII[20] = {-1, 0, 1, 1, 1, 0, -1, -1, -1, 1, 1, -1, -1, 0, 1, 1, 1, 0, -1, -1};
static const double ETI[20] = {-1, -1, -1, 0, 1, 1, 1, 0, -1, -1, 1, 1, -1, -1, -1, 0, 1, 1, 1, 0};
static const double ZEI[20] = {-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1};
static const int KIND[20] = {1, 2, 1, 3, 1, 2, 1, 3, 4, 4, 4, 4, 1, 2, 1, 3, 1, 2, 1, 3};

/* ihex.f:2433-2545 */
static void dn2031(double *DNXI, double *DNET, double *DNZE, double XI, double ET, double ZE)
{
  for (int I = 0; I < 20; I++) {
    switch (KIND[I]) {
    case 1:
      DNXI[I] = .125 * XII[I] * (1. + ET * ETI[I]) * (1. + ZE * ZEI[I]) *
                (2. * XI * XII[I] + ET * ETI[I] + ZE * ZEI[I] - 1.);
      DNET[I] = .125 * ETI[I] * (1. + XI * XII[I]) * (1. + ZE * ZEI[I]) *
                (XI * XII[I] + 2. * ET * ETI[I] + ZE * ZEI[I] - 1.);
      DNZE[I] = .125 * ZEI[I] * (1. + XI * XII[I]) * (1. + ET * ETI[I]) *
                (XI * XII[I] + ET * ETI[I] + 2. * ZE * ZEI[I] - 1.);
      break;
    case 2:
      DNXI[I] = -.5 * XI * (1. + ET * ETI[I]) * (1. + ZE * ZEI[I]);
      DNET[I] = .25 * ETI[I] * (1. - XI * XI) * (1. + ZE * ZEI[I]);
      DNZE[I] = .25 * ZEI[I] * (1. - XI * XI) * (1. + ET * ETI[I]);
      break;
    case 3:
      DNXI[I] = .25 * XII[I] * (1. - ET * ET) * (1. + ZE * ZEI[I]);
      DNET[I] = -.5 * ET * (1. + XI * XII[I]) * (1. + ZE * ZEI[I]);
      DNZE[I] = .25 * ZEI[I] * (1. + XI * XII[I]) * (1. - ET * ET);
      break;
    default:
      DNXI[I] = .25 * XII[I] * (1. + ET * ETI[I]) * (1. - ZE * ZE);
      DNET[I] = .25 * ETI[I] * (1. + XI * XII[I]) * (1. - ZE * ZE);
      DNZE[I] = -.5 * ZE * (1. + XI * XII[I]) * (1. + ET * ETI[I]);
    }
  }
}

/* jaci31.f with MEK = 20 */
static int jaci20(double JI[3][3], const double *dxi, const double *det_, const double *dze,
                  const double *XG, const double *YG, const double *ZG)
{
  const double EPS = DBL_MIN * 100.0;
  double J[3][3], DETJ;
  memset(J, 0, sizeof(J));
  for (int i = 0; i < 20; i++) {
    J[0][0] += dxi[i] * XG[i];  J[0][1] += dxi[i] * YG[i];  J[0][2] += dxi[i] * ZG[i];
    J[1][0] += det_[i] * XG[i]; J[1][1] += det_[i] * YG[i]; J[1][2] += det_[i] * ZG[i];
    J[2][0] += dze[i] * XG[i];  J[2][1] += dze[i] * YG[i];  J[2][2] += dze[i] * ZG[i];
  }
  DETJ = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) +
         J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  if (fabs(DETJ) - EPS <= 0.0) return -1;
  JI[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / DETJ;
  JI[0][1] = (J[2][1] * J[0][2] - J[2][2] * J[0][1]) / DETJ;
  JI[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / DETJ;
  JI[1][0] = (J[2][0] * J[1][2] - J[2][2] * J[1][0]) / DETJ;
  JI[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / DETJ;
  JI[1][2] = (J[1][0] * J[0][2] - J[1][2] * J[0][0]) / DETJ;
  JI[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / DETJ;
  JI[2][1] = (J[2][0] * J[0][1] - J[2][1] * J[0][0]) / DETJ;
  JI[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / DETJ;
  return 0;
}

/* ihex.f:224-560 with IOPXP = IOPE0 = 0, NSTRXI = NSTRET = NSTRZE = n (0 -> the 27-point nodal
 * grid, 2 -> 2x2x2 Gauss points).  SIG(6,27) column-major. */
static int ihex32(double *SIG, const double *V, const double *XG, const double *YG,
                  const double *ZG, double YOUNG, double RNY, int n)
{
  double DNXI[20], DNET[20], DNZE[20], JI[3][3], B[3], DB[20][6][3], XI[4], D, D1, D2;
  int NXI, N = 0;
  memset(SIG, 0, sizeof(double) * 6 * 27);
  D = YOUNG * (1. - RNY) / ((1. + RNY) * (1. - 2. * RNY));
  D1 = D * RNY / (1. - RNY);
  D2 = D * (1. - 2. * RNY) / (2. * (1. - RNY));
  if (n == 2) {
    NXI = 2;
    XI[0] = -(double)(float).577350269189626;
    XI[1] = (double)(float).577350269189626;
  } else if (n == 0) {
    NXI = 3;
    for (int i = 1; i <= 3; i++) XI[i - 1] = i - 2.0;
  } else
    return -1;
  for (int M = 0; M < NXI; M++)
    for (int L = 0; L < NXI; L++)
      for (int K = 0; K < NXI; K++) {
        dn2031(DNXI, DNET, DNZE, XI[K], XI[L], XI[M]);
        if (jaci20(JI, DNXI, DNET, DNZE, XG, YG, ZG) != 0) return -1;
        for (int J = 0; J < 20; J++) {
          B[0] = JI[0][0] * DNXI[J] + JI[0][1] * DNET[J] + JI[0][2] * DNZE[J];
          B[1] = JI[1][0] * DNXI[J] + JI[1][1] * DNET[J] + JI[1][2] * DNZE[J];
          B[2] = JI[2][0] * DNXI[J] + JI[2][1] * DNET[J] + JI[2][2] * DNZE[J];
          DB[J][0][0] = D * B[0];
          DB[J][1][0] = D1 * B[0];
          DB[J][2][0] = DB[J][1][0];
          DB[J][3][0] = D2 * B[1];
          DB[J][4][0] = D2 * B[2];
          DB[J][5][0] = 0.0;
          DB[J][0][1] = D1 * B[1];
          DB[J][1][1] = D * B[1];
          DB[J][2][1] = DB[J][0][1];
          DB[J][3][1] = D2 * B[0];
          DB[J][4][1] = 0.0;
          DB[J][5][1] = DB[J][4][0];
          DB[J][0][2] = D1 * B[2];
          DB[J][1][2] = DB[J][0][2];
          DB[J][2][2] = D * B[2];
          DB[J][3][2] = 0.0;
          DB[J][4][2] = DB[J][3][1];
          DB[J][5][2] = DB[J][3][0];
        }
        for (int I = 0; I < 20; I++)
          for (int J = 0; J < 6; J++)
            SIG[J + 6 * N] = SIG[J + 6 * N] + DB[I][J][0] * V[3 * I] + DB[I][J][1] * V[3 * I + 1] +
                             DB[I][J][2] * V[3 * I + 2];
        N++;
      }
  return 0;
}

/* elStressModule.f90:1472-1584; sigma(6,20), epsil(6,20) column-major */
int orc_str43(const double xg[20], const double yg[20], const double zg[20], double emod,
              double rny, int stressForm, const double v[60], double sigma[120],
              double epsil[120])
{
  /* SIGG column (1-based) of the 27-point grid that holds node n (elStressModule.f90:1541-1560) */
  static const int NODE_PT[20] = {1, 2, 3, 6, 9, 8, 7, 4, 10, 12, 18, 16, 19, 20, 21, 24, 27, 26, 25, 22};
  const double one_p = sqrt(3.0);
  double SIGG[6 * 27], Einv[36];
  memset(Einv, 0, sizeof(Einv));
  Einv[0] = 1.0 / emod;
  Einv[1] = -rny / emod;
  Einv[2] = Einv[1];
  Einv[6] = Einv[1];  Einv[7] = Einv[0];  Einv[8] = Einv[1];
  Einv[12] = Einv[1]; Einv[13] = Einv[1]; Einv[14] = Einv[0];
  Einv[21] = 2.0 * (1.0 + rny) / emod;
  Einv[28] = Einv[21];
  Einv[35] = Einv[21];
  if (ihex32(SIGG, v, xg, yg, zg, emod, rny, stressForm == 0 ? 0 : 2) != 0) return 1;
  if (stressForm == 0) {
    for (int n = 0; n < 20; n++)
      for (int c = 0; c < 6; c++) sigma[c + 6 * n] = SIGG[c + 6 * (NODE_PT[n] - 1)];
  } else {
    memset(sigma, 0, sizeof(double) * 120);
    for (int n = 0; n < 20; n++) {
      int l = 0;
      for (int k = -1; k <= 1; k += 2) {
        double z = 1.0 + k * (one_p * ZEI[n]);
        for (int j = -1; j <= 1; j += 2) {
          double y = 1.0 + j * (one_p * ETI[n]);
          for (int i = -1; i <= 1; i += 2) {
            double x = 1.0 + i * (one_p * XII[n]);
            for (int c = 0; c < 6; c++)
              sigma[c + 6 * n] = sigma[c + 6 * n] + SIGG[c + 6 * l] * x * y * z * 0.125;
            l++;
          }
        }
      }
    }
  }
  for (int p = 0; p < 20; p++)
    for (int r = 0; r < 6; r++) {
      double s = 0.0;
      for (int k = 0; k < 6; k++) s += Einv[r + 6 * k] * sigma[k + 6 * p];
      epsil[r + 6 * p] = s;
    }
  return 0;
}
