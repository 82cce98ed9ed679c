/* tet10.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): 10-node tetrahedron STR41.
 * Follows src/vpmStress/elStressModule.f90:1308-1375 (STR41), src/Femlib/itet.f:7-87 (DN1031),
 * :701-997 (ITET32), src/Femlib/jaci31.f:7-133 (JACO31, JACI31), isoMatModule.f90:95-120.
 * F77 literals without D exponent are REAL*4 promoted to double (Femlib is compiled without
 * -fdefault-real-8): .585410196625 / .138196601125 (itet.f:837-838) are rounded to float. */
#include "oracle.h"
#include <math.h>
#include <string.h>

/* itet.f:48-84 */
static void dn1031(double d1[10], double d2[10], double d3[10], double RL1, double RL2,
                   double RL3, double RL4)
{
  d1[0] = 4. * RL1 - 1.;  d1[1] = 4. * RL2;  d1[2] = 0.;  d1[3] = 0.;  d1[4] = 0.;
  d1[5] = 4. * RL3;  d1[6] = 4. * (RL4 - RL1);  d1[7] = -4. * RL2;  d1[8] = -4. * RL3;
  d1[9] = -4. * RL4 + 1.;
  d2[0] = 0.;  d2[1] = 4. * RL1;  d2[2] = 4. * RL2 - 1.;  d2[3] = 4. * RL3;  d2[4] = 0.;
  d2[5] = 0.;  d2[6] = -4. * RL1;  d2[7] = 4. * (RL4 - RL2);  d2[8] = -4. * RL3;
  d2[9] = -4. * RL4 + 1.;
  d3[0] = 0.;  d3[1] = 0.;  d3[2] = 0.;  d3[3] = 4. * RL2;  d3[4] = 4. * RL3 - 1.;
  d3[5] = 4. * RL1;  d3[6] = -4. * RL1;  d3[7] = -4. * RL2;  d3[8] = 4. * RL4 - 4. * RL3;
  d3[9] = -4. * RL4 + 1.;
}

/* jaci31.f: JACO31 + JACI31.  JI row-major here: JI[i][j] = JI(i+1,j+1). */
static int jaci31(double JI[3][3], const double *dxi, const double *det_, const double *dze,
                  const double *XG, const double *YG, const double *ZG, int mek)
{
  const double EPS = DBL_MIN * 100.0; /* tiny(1.0D0)*100.0D0 */
  double J[3][3], DETJ;
  memset(J, 0, sizeof(J));
  for (int i = 0; i < mek; i++) {
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

/* itet.f:701-997 with IOPXP = IOPE0 = 0.  SIG(6,NSTRP) column-major, V(3,10). */
static int itet32(double *SIG, const double *V, const double *XG, const double *YG,
                  const double *ZG, double YOUNG, double RNY, int NSTRP)
{
  enum { MEK = 10 };
  double RL1[10], RL2[10], RL3[10], RL4[10], DNL1[10], DNL2[10], DNL3[10];
  double JI[3][3], B[3], DB[10][6][3], D, D1, D2, ALFA, BETA;

  for (int i = 0; i < 6 * NSTRP; i++) SIG[i] = 0.0;
  for (int i = 0; i < 10; i++) RL1[i] = RL2[i] = RL3[i] = RL4[i] = 0.0;

  D = YOUNG * (1. - RNY) / ((1. + RNY) * (1. - 2. * RNY));
  D1 = D * RNY / (1. - RNY);
  D2 = D * (1. - 2. * RNY) / (2. * (1. - RNY));

  switch (NSTRP) {
  case 1:
    RL1[0] = .25; RL2[0] = .25; RL3[0] = .25;
    break;
  case 4:
    ALFA = (double)(float).585410196625; /* REAL*4 literal, itet.f:837 */
    BETA = (double)(float).138196601125; /* REAL*4 literal, itet.f:838 */
    RL1[0] = ALFA; RL2[0] = BETA; RL3[0] = BETA;
    RL1[1] = BETA; RL2[1] = ALFA; RL3[1] = BETA;
    RL1[2] = BETA; RL2[2] = BETA; RL3[2] = ALFA;
    RL1[3] = BETA; RL2[3] = BETA; RL3[3] = BETA;
    break;
  case 10:
    RL1[0] = 1.0;
    RL1[1] = 0.5; RL2[1] = 0.5;
    RL2[2] = 1.0;
    RL2[3] = 0.5; RL3[3] = 0.5;
    RL3[4] = 1.0;
    RL1[5] = 0.5; RL3[5] = 0.5;
    RL1[6] = 0.5;
    RL2[7] = 0.5;
    RL3[8] = 0.5;
    break;
  default:
    return -1;
  }

  for (int L = 0; L < NSTRP; L++) {
    RL4[L] = 1.0 - RL1[L] - RL2[L] - RL3[L];
    dn1031(DNL1, DNL2, DNL3, RL1[L], RL2[L], RL3[L], RL4[L]);
    if (jaci31(JI, DNL1, DNL2, DNL3, XG, YG, ZG, MEK) < 0) return -1;
    for (int J = 0; J < MEK; J++) {
      B[0] = JI[0][0] * DNL1[J] + JI[0][1] * DNL2[J] + JI[0][2] * DNL3[J];
      B[1] = JI[1][0] * DNL1[J] + JI[1][1] * DNL2[J] + JI[1][2] * DNL3[J];
      B[2] = JI[2][0] * DNL1[J] + JI[2][1] * DNL2[J] + JI[2][2] * DNL3[J];
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
    for (int I = 0; I < MEK; I++)
      for (int J = 0; J < 6; J++)
        SIG[J + 6 * L] = SIG[J + 6 * L] + DB[I][J][0] * V[3 * I] + DB[I][J][1] * V[3 * I + 1] +
                         DB[I][J][2] * V[3 * I + 2];
  }
  return 0;
}

/* elStressModule.f90:1308-1375; sigma(6,10), epsil(6,10) column-major */
int orc_str41(const double xg[10], const double yg[10], const double zg[10], double emod,
              double rny, int stressForm, const double v[30], double sigma[60],
              double epsil[60])
{
  const double alpha_p = 1.927051062810166, beta_p = -0.309017015969668;
  double SIGG[60], Einv[36];
  int n = stressForm == 0 ? 10 : 4;

  /* isoMat3Dinv, isoMatModule.f90:95-120 */
  memset(Einv, 0, sizeof(Einv));
  Einv[0] = 1.0 / emod;
  Einv[1] = -rny / emod;
  Einv[2] = Einv[1];
  Einv[6] = Einv[1];  Einv[7] = Einv[0];  Einv[8] = Einv[1];
  Einv[12] = Einv[1]; Einv[13] = Einv[1]; Einv[14] = Einv[0];
  Einv[21] = 2.0 * (1.0 + rny) / emod;
  Einv[28] = Einv[21];
  Einv[35] = Einv[21];

  if (itet32(SIGG, v, xg, yg, zg, emod, rny, n) != 0) return 1;

  if (stressForm == 0)
    memcpy(sigma, SIGG, sizeof(double) * 60);
  else {
#define SG(c, p) SIGG[(c) + 6 * ((p)-1)]
#define SI(c, p) sigma[(c) + 6 * ((p)-1)]
    for (int c = 0; c < 6; c++) {
      SI(c, 1) = alpha_p * SG(c, 1) + beta_p * (SG(c, 2) + SG(c, 3) + SG(c, 4));
      SI(c, 3) = alpha_p * SG(c, 2) + beta_p * (SG(c, 1) + SG(c, 3) + SG(c, 4));
      SI(c, 5) = alpha_p * SG(c, 3) + beta_p * (SG(c, 1) + SG(c, 2) + SG(c, 4));
      SI(c, 10) = alpha_p * SG(c, 4) + beta_p * (SG(c, 1) + SG(c, 2) + SG(c, 3));
      SI(c, 2) = 0.5 * (SI(c, 1) + SI(c, 3));
      SI(c, 4) = 0.5 * (SI(c, 3) + SI(c, 5));
      SI(c, 6) = 0.5 * (SI(c, 5) + SI(c, 1));
      SI(c, 7) = 0.5 * (SI(c, 1) + SI(c, 10));
      SI(c, 8) = 0.5 * (SI(c, 3) + SI(c, 10));
      SI(c, 9) = 0.5 * (SI(c, 5) + SI(c, 10));
    }
#undef SG
#undef SI
  }

  for (int p = 0; p < 10; p++)
    for (int r = 0; r < 6; r++) {
      double s = 0.0;
      for (int k = 0; k < 6; k++) s += Einv[r + 6 * k] * sigma[k + 6 * p];
      epsil[r + 6 * p] = s;
    }
  return 0;
}
