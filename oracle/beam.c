/* beam.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): 2-node beam section forces STR11.
 * Follows src/vpmStress/elStressModule.f90:402-515 (STR11), src/Femlib/beam.f:11-120 (BEAM31),
 * :619-770 (BELS31 local Timoshenko stiffness incl. shear-centre coupling),
 * src/Femlib/beamaux.f:48-105 (DCOS30), :106-165 (MPRO30), :166-208 (TRIX30),
 * SAM/src/Mat.f:1170-1244 (VECTRA), :1245-1330 (MATTRA).
 * beam[] layout (ORC_NBEAM doubles per element, what ffl_getcoor / ffl_getbeamsection /
 * ffl_getpinflags return, FFlLinkHandler_F.C:745-812,1131-1193):
 *   [0:5) X(1:5), [5:10) Y(1:5), [10:15) Z(1:5)  (1,2 = beam ends incl. eccentricity, 3 = point on
 *   the local Z axis, 4,5 = the nodes), [15:29) BSEC(1:14), [29] IPA, [30] IPB. */
#include "oracle.h"
#include <math.h>
#include <string.h>

#define EKm(i, j) EK[((i)-1) + 12 * ((j)-1)]

/* beamaux.f:48-105; C column-major 3x3 */
static void dcos30(double *C, const double *X, const double *Y, const double *Z)
{
#define Cm(i, j) C[((i)-1) + 3 * ((j)-1)]
  double CX, CY, CZ, AB;
  CX = X[1] - X[0]; CY = Y[1] - Y[0]; CZ = Z[1] - Z[0];
  AB = sqrt(CX * CX + CY * CY + CZ * CZ);
  Cm(1, 1) = CX / AB; Cm(1, 2) = CY / AB; Cm(1, 3) = CZ / AB;
  CX = Cm(1, 3) * (Y[2] - Y[0]) - Cm(1, 2) * (Z[2] - Z[0]);
  CY = Cm(1, 1) * (Z[2] - Z[0]) - Cm(1, 3) * (X[2] - X[0]);
  CZ = Cm(1, 2) * (X[2] - X[0]) - Cm(1, 1) * (Y[2] - Y[0]);
  AB = sqrt(CX * CX + CY * CY + CZ * CZ);
  Cm(2, 1) = CX / AB; Cm(2, 2) = CY / AB; Cm(2, 3) = CZ / AB;
  CX = Cm(1, 2) * Cm(2, 3) - Cm(1, 3) * Cm(2, 2);
  CY = Cm(1, 3) * Cm(2, 1) - Cm(1, 1) * Cm(2, 3);
  CZ = Cm(1, 1) * Cm(2, 2) - Cm(1, 2) * Cm(2, 1);
  AB = sqrt(CX * CX + CY * CY + CZ * CZ);
  Cm(3, 1) = CX / AB; Cm(3, 2) = CY / AB; Cm(3, 3) = CZ / AB;
}

/* beam.f:619-770 */
static void bels31(double *EK, double BL, const double *EP, const double *CA, const double *XS)
{
  double E = EP[0], G = EP[1], A = EP[2], RIY = EP[3], RIZ = EP[4], RIT = EP[5];
  double CAY = CA[0], CAZ = CA[1], YS = XS[0], ZS = XS[1];
  double EA, EIY, EIZ, ALY, ALZ, GIT, BE, BA;
  if (E < -1.0e-16) {
    EA = A; EIY = RIY; EIZ = RIZ;
    ALY = CAY > 1.0e-16 ? 12.0 * EIY / (CAY * BL * BL) : 0.0;
    ALZ = CAZ > 1.0e-16 ? 12.0 * EIZ / (CAZ * BL * BL) : 0.0;
    GIT = RIT;
  } else {
    EA = E * A; EIY = E * RIY; EIZ = E * RIZ;
    ALY = 12.0 * CAY * EIY / (A * G * BL * BL);
    ALZ = 12.0 * CAZ * EIZ / (A * G * BL * BL);
    GIT = G * RIT;
  }
  memset(EK, 0, sizeof(double) * 144);
  EKm(1, 1) = EA / BL;
  EKm(2, 2) = 12. * EIY / (BL * BL * BL * (1. + ALY));
  EKm(3, 3) = 12. * EIZ / (BL * BL * BL * (1. + ALZ));
  EKm(4, 4) = GIT / BL;
  EKm(3, 5) = -.5 * BL * EKm(3, 3);
  EKm(2, 6) = .5 * BL * EKm(2, 2);
  EKm(5, 5) = EIZ * (4. + ALZ) / (BL * (1. + ALZ));
  EKm(6, 6) = EIY * (4. + ALY) / (BL * (1. + ALY));
  EKm(1, 7) = -EKm(1, 1);
  EKm(7, 7) = EKm(1, 1);
  EKm(2, 8) = -EKm(2, 2);
  EKm(6, 8) = -EKm(2, 6);
  EKm(8, 8) = EKm(2, 2);
  EKm(3, 9) = -EKm(3, 3);
  EKm(5, 9) = -EKm(3, 5);
  EKm(9, 9) = EKm(3, 3);
  EKm(4, 10) = -EKm(4, 4);
  EKm(10, 10) = EKm(4, 4);
  EKm(3, 11) = EKm(3, 5);
  EKm(5, 11) = EIZ * (2. - ALZ) / (BL * (1. + ALZ));
  EKm(9, 11) = -EKm(3, 5);
  EKm(11, 11) = EKm(5, 5);
  EKm(2, 12) = EKm(2, 6);
  EKm(6, 12) = EIY * (2. - ALY) / (BL * (1. + ALY));
  EKm(8, 12) = -EKm(2, 6);
  EKm(12, 12) = EKm(6, 6);
  for (int i = 1; i <= 12; i++)
    for (int j = 1; j <= i; j++) EKm(i, j) = EKm(j, i);
  BE = fabs(YS) + fabs(ZS);
  BA = sqrt(A);
  BA = (double)(float).00001 * BA; /* REAL*4 literal .00001 (threshold only) */
  if (BA - BE < 0.0) {
    for (int i = 1; i <= 12; i++) {
      EKm(4, i) = EKm(4, i) - ZS * EKm(2, i) + YS * EKm(3, i);
      EKm(10, i) = EKm(10, i) - ZS * EKm(8, i) + YS * EKm(9, i);
    }
    for (int i = 1; i <= 12; i++) {
      EKm(i, 4) = EKm(i, 4) - EKm(i, 2) * ZS + EKm(i, 3) * YS;
      EKm(i, 10) = EKm(i, 10) - EKm(i, 8) * ZS + EKm(i, 9) * YS;
    }
  }
}

/* A = TT*A*T with the 3x3 block T on the diagonal at position K (SAM MATTRA, Mat.f:1245-1330) */
static void mattra(const double *T, double *EK, int K)
{
  double F[144], tmp[144];
  memset(F, 0, sizeof(F));
  for (int i = 0; i < 12; i++) F[i + 12 * i] = 1.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) F[(K - 1 + i) + 12 * (K - 1 + j)] = T[i + 3 * j];
  for (int i = 0; i < 12; i++)
    for (int j = 0; j < 12; j++) {
      double s = 0.0;
      for (int k = 0; k < 12; k++) s += F[k + 12 * i] * EK[k + 12 * j];
      tmp[i + 12 * j] = s;
    }
  for (int i = 0; i < 12; i++)
    for (int j = 0; j < 12; j++) {
      double s = 0.0;
      for (int k = 0; k < 12; k++) s += tmp[i + 12 * k] * F[k + 12 * j];
      EK[i + 12 * j] = s;
    }
}

/* beamaux.f:106-165: EK = blockdiag(C)^T EK blockdiag(C), then symmetrised from the lower triangle */
static void mpro30(double *EK, const double *C)
{
#define Cm(i, j) C[((i)-1) + 3 * ((j)-1)]
  double B[3][3];
  for (int I = 0; I <= 9; I += 3)
    for (int J = 0; J <= 9; J += 3) {
      memset(B, 0, sizeof(B));
      for (int IA = 1; IA <= 3; IA++)
        for (int KA = 1; KA <= 3; KA++)
          for (int JA = 1; JA <= 3; JA++) B[JA - 1][IA - 1] += Cm(KA, JA) * EKm(J + KA, I + IA);
      for (int IA = 1; IA <= 3; IA++)
        for (int KA = 1; KA <= 3; KA++) EKm(J + KA, I + IA) = 0.0;
      for (int IA = 1; IA <= 3; IA++)
        for (int KA = 1; KA <= 3; KA++)
          for (int JA = 1; JA <= 3; JA++) EKm(J + KA, I + IA) += B[KA - 1][JA - 1] * Cm(JA, IA);
    }
  for (int I = 1; I <= 12; I++)
    for (int J = I; J <= 12; J++) EKm(I, J) = EKm(J, I);
#undef Cm
}

/* beamaux.f:166-208 */
static void trix30(double *EK, double X1, double Y1, double Z1, double X2, double Y2, double Z2)
{
  for (int i = 1; i <= 12; i++) {
    EKm(4, i) = EKm(4, i) + Z1 * EKm(2, i) - Y1 * EKm(3, i);
    EKm(5, i) = EKm(5, i) - Z1 * EKm(1, i) + X1 * EKm(3, i);
    EKm(6, i) = EKm(6, i) + Y1 * EKm(1, i) - X1 * EKm(2, i);
    EKm(10, i) = EKm(10, i) + Z2 * EKm(8, i) - Y2 * EKm(9, i);
    EKm(11, i) = EKm(11, i) - Z2 * EKm(7, i) + X2 * EKm(9, i);
    EKm(12, i) = EKm(12, i) + Y2 * EKm(7, i) - X2 * EKm(8, i);
  }
  for (int i = 1; i <= 12; i++) {
    EKm(i, 4) = EKm(i, 4) + Z1 * EKm(i, 2) - Y1 * EKm(i, 3);
    EKm(i, 5) = EKm(i, 5) - Z1 * EKm(i, 1) + X1 * EKm(i, 3);
    EKm(i, 6) = EKm(i, 6) + Y1 * EKm(i, 1) - X1 * EKm(i, 2);
    EKm(i, 10) = EKm(i, 10) + Z2 * EKm(i, 8) - Y2 * EKm(i, 9);
    EKm(i, 11) = EKm(i, 11) - Z2 * EKm(i, 7) + X2 * EKm(i, 9);
    EKm(i, 12) = EKm(i, 12) + Y2 * EKm(i, 7) - X2 * EKm(i, 8);
  }
}

/* beam.f:11-120.  XS is updated in place when PHI rotates the section axes, like the reference. */
static int beam31(double *EK, const double *X, const double *Y, const double *Z, const double *EP,
                  const double *CA, double *XS, double EFFLEN, double PHI, int IPINA, int IPINB)
{
  double T2[9], E1[3], E2[3], BL, BA, BX, BY, BZ;
#define T2m(i, j) T2[((i)-1) + 3 * ((j)-1)]
  if (EFFLEN > 0.0)
    BL = EFFLEN;
  else {
    BX = X[1] - X[0]; BY = Y[1] - Y[0]; BZ = Z[1] - Z[0];
    BL = sqrt(BX * BX + BY * BY + BZ * BZ);
  }
  BA = fabs(X[0]) + fabs(X[1]) + fabs(Y[0]) + fabs(Y[1]) + fabs(Z[0]) + fabs(Z[1]);
  BA = 1.0e-6 * BA;
  if (BL - BA <= 0.0) return -1;
  BX = X[2] - X[0]; BY = Y[2] - Y[0]; BZ = Z[2] - Z[0];
  BZ = sqrt(BX * BX + BY * BY + BZ * BZ);
  if (BZ - BA <= 0.0) return -2;
  if (EP[1] <= 1.0e-16 || EP[2] <= 1.0e-16) return -3;
  dcos30(T2, X, Y, Z);
  if (fabs(PHI) > 1.0e-6) {
    double fi = PHI * atan(1.0) / 4.5e1, cf = cos(fi), sf = sin(fi), a, b;
    for (int i = 1; i <= 3; i++) {
      a = cf * T2m(2, i) + sf * T2m(3, i);
      b = cf * T2m(3, i) - sf * T2m(2, i);
      T2m(2, i) = a;
      T2m(3, i) = b;
    }
    a = cf * XS[0] + sf * XS[1];
    b = cf * XS[1] - sf * XS[0];
    XS[0] = a;
    XS[1] = b;
  }
  bels31(EK, BL, EP, CA, XS);
  if (IPINA <= 0 && IPINB <= 0)
    mpro30(EK, T2);
  else if (IPINA <= 0) {
    mattra(T2, EK, 1);
    mattra(T2, EK, 4);
  } else if (IPINB <= 0) {
    mattra(T2, EK, 7);
    mattra(T2, EK, 10);
  }
  E1[0] = X[3] - X[0]; E1[1] = Y[3] - Y[0]; E1[2] = Z[3] - Z[0];
  E2[0] = X[4] - X[1]; E2[1] = Y[4] - Y[1]; E2[2] = Z[4] - Z[1];
  if (IPINA > 0) { /* VECTRA(T2,E1,...,IFLAG=1): E1 = T2*E1 */
    double w[3];
    for (int i = 1; i <= 3; i++) w[i - 1] = T2m(i, 1) * E1[0] + T2m(i, 2) * E1[1] + T2m(i, 3) * E1[2];
    memcpy(E1, w, sizeof(w));
  }
  if (IPINB > 0) {
    double w[3];
    for (int i = 1; i <= 3; i++) w[i - 1] = T2m(i, 1) * E2[0] + T2m(i, 2) * E2[1] + T2m(i, 3) * E2[2];
    memcpy(E2, w, sizeof(w));
  }
  trix30(EK, E1[0], E1[1], E1[2], E2[0], E2[1], E2[2]);
  return 0;
#undef T2m
}

/* elStressModule.f90:402-515.  SF(6,2) column-major.  Returns 0, or 1 if BEAM31 failed. */
int orc_str11(const double *beam, const double ev[12], double SF[12])
{
  double XG[5], YG[5], ZG[5], BSEC[14], EK[144], Sg[6], T[9], SN[3], SM[3], Ex, Ey, Ez, Bl;
  int IPA, IPB;
  memcpy(XG, beam, sizeof(XG));
  memcpy(YG, beam + 5, sizeof(YG));
  memcpy(ZG, beam + 10, sizeof(ZG));
  memcpy(BSEC, beam + 15, sizeof(BSEC));
  IPA = (int)beam[29];
  IPB = (int)beam[30];
  for (int i = 0; i < 12; i++) SF[i] = 0.0;
  /* BEAM31(EK,XG,YG,ZG,BSEC(2),BSEC(9),BSEC(11),BSEC(13),BSEC(14),IPA,IPB,...) */
  if (beam31(EK, XG, YG, ZG, &BSEC[1], &BSEC[8], &BSEC[10], BSEC[12], BSEC[13], IPA, IPB) != 0)
    return 1;
  for (int i = 1; i <= 6; i++) {
    double s = 0.0;
    for (int j = 1; j <= 12; j++) s += EKm(i, j) * ev[j - 1];
    Sg[i - 1] = s;
  }
  Ex = XG[3] - XG[0]; Ey = YG[3] - YG[0]; Ez = ZG[3] - ZG[0];
  Sg[3] = Sg[3] - Ez * Sg[1] + Ey * Sg[2];
  Sg[4] = Sg[4] + Ez * Sg[0] - Ex * Sg[2];
  Sg[5] = Sg[5] - Ey * Sg[0] + Ex * Sg[1];
  dcos30(T, XG, YG, ZG);
  for (int i = 0; i < 3; i++) {
    SN[i] = T[i] * Sg[0] + T[i + 3] * Sg[1] + T[i + 6] * Sg[2];
    SM[i] = T[i] * Sg[3] + T[i + 3] * Sg[4] + T[i + 6] * Sg[5];
  }
  SF[0] = -SN[0];
  SF[1] = SN[1];
  SF[2] = SN[2];
  SF[3] = -SM[0] + BSEC[9] * SN[2] - BSEC[10] * SN[1];
  SF[4] = SM[1];
  SF[5] = SM[2];
  Ex = XG[1] - XG[0]; Ey = YG[1] - YG[0]; Ez = ZG[1] - ZG[0];
  Bl = sqrt(Ex * Ex + Ey * Ey + Ez * Ez);
  for (int i = 0; i < 6; i++) SF[6 + i] = SF[i];
  SF[6 + 4] = SF[6 + 4] + SN[2] * Bl;
  SF[6 + 5] = SF[6 + 5] + SN[1] * Bl;
  return 0;
}
