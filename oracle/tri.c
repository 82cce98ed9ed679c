/* tri.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): ANDES/FTS triangle STR23.
 * Follows src/vpmStress/elStressModule.f90:901-999 (STR23), src/Femlib/ftsa.f:7-260
 * (FTSA31 local coordinates, FTSA32), src/Femlib/hlst.f:7-355 (HLST31 kappa matrix, HLST32),
 * src/Femlib/nyteba.f:7-364 (TEBA31 kappa matrix, TEBA32), src/Femlib/fts.f:446-490 (FTS38),
 * src/Femlib/beamaux.f:209-260 (DIRC30), src/Femlib/dinv12.f (DGETRF+DGETRI inverse).
 *
 * Precision traps kept on purpose (Femlib is compiled without -fdefault-real-8, so literals
 * without a D exponent are REAL*4): the 7-point rule in TEBA31 (nyteba.f:35-43) and
 * ZZ = 1./3. in FTSA32 (ftsa.f:223) are float values promoted to double.
 * LAPACK is an un-vendored system dependency of the reference (find_package(LAPACK),
 * src/Femlib/CMakeLists.txt:27, version unpinned); DGETRF/DGETRI are restated here as LU with
 * partial pivoting.  Pivot-order rounding differences are O(1e-15) relative. */
#include "oracle.h"
#include <math.h>
#include <string.h>

#define F32(x) ((double)(float)(x))

/* In-place inverse of the column-major n x n matrix a (n <= 9): LU with partial pivoting
 * (DGETRF), then inverse from the factors (DGETRI).  Returns nonzero if singular. */
static int dinv12(int n, double *a)
{
  int piv[9];
  double inv[81], col[9];
#define A(i, j) a[(i) + n * (j)]
  for (int k = 0; k < n; k++) {
    int p = k;
    double big = fabs(A(k, k));
    for (int i = k + 1; i < n; i++)
      if (fabs(A(i, k)) > big) { big = fabs(A(i, k)); p = i; }
    piv[k] = p;
    if (A(p, k) == 0.0) return k + 1;
    if (p != k)
      for (int j = 0; j < n; j++) { double t = A(k, j); A(k, j) = A(p, j); A(p, j) = t; }
    for (int i = k + 1; i < n; i++) A(i, k) = A(i, k) / A(k, k);
    for (int j = k + 1; j < n; j++)
      for (int i = k + 1; i < n; i++) A(i, j) = A(i, j) - A(i, k) * A(k, j);
  }
  /* Solve A * X = I column by column using P, L, U */
  for (int c = 0; c < n; c++) {
    for (int i = 0; i < n; i++) col[i] = (i == c) ? 1.0 : 0.0;
    for (int k = 0; k < n; k++)
      if (piv[k] != k) { double t = col[k]; col[k] = col[piv[k]]; col[piv[k]] = t; }
    for (int i = 0; i < n; i++)
      for (int k = 0; k < i; k++) col[i] -= A(i, k) * col[k];
    for (int i = n - 1; i >= 0; i--) {
      for (int k = i + 1; k < n; k++) col[i] -= A(i, k) * col[k];
      col[i] = col[i] / A(i, i);
    }
    for (int i = 0; i < n; i++) inv[i + n * c] = col[i];
  }
  memcpy(a, inv, sizeof(double) * n * n);
#undef A
  return 0;
}

/* beamaux.f:209-260; C column-major 3x3 */
static void dirc30(double *C, const double *X, const double *Y, const double *Z)
{
#define Cm(i, j) C[((i)-1) + 3 * ((j)-1)]
  double CX, CY, CZ, AB;
  CX = X[1] - X[0]; CY = Y[1] - Y[0]; CZ = Z[1] - Z[0];
  AB = sqrt(CX * CX + CY * CY + CZ * CZ);
  Cm(1, 1) = CX / AB; Cm(1, 2) = CY / AB; Cm(1, 3) = CZ / AB;
  CX = Cm(1, 2) * (Z[2] - Z[0]) - Cm(1, 3) * (Y[2] - Y[0]);
  CY = Cm(1, 3) * (X[2] - X[0]) - Cm(1, 1) * (Z[2] - Z[0]);
  CZ = Cm(1, 1) * (Y[2] - Y[0]) - Cm(1, 2) * (X[2] - X[0]);
  AB = sqrt(CX * CX + CY * CY + CZ * CZ);
  Cm(3, 1) = CX / AB; Cm(3, 2) = CY / AB; Cm(3, 3) = CZ / AB;
  CX = Cm(3, 2) * Cm(1, 3) - Cm(3, 3) * Cm(1, 2);
  CY = Cm(3, 3) * Cm(1, 1) - Cm(3, 1) * Cm(1, 3);
  CZ = Cm(3, 1) * Cm(1, 2) - Cm(3, 2) * Cm(1, 1);
  AB = sqrt(CX * CX + CY * CY + CZ * CZ);
  Cm(2, 1) = CX / AB; Cm(2, 2) = CY / AB; Cm(2, 3) = CZ / AB;
#undef Cm
}

/* ftsa.f:60-80 and :224-242: local in-plane coordinates of the three corners */
static void local_xy(const double *X, const double *Y, const double *Z, double XL[3],
                     double YL[3])
{
  double X21 = X[1] - X[0], Y21 = Y[1] - Y[0], Z21 = Z[1] - Z[0];
  double X31 = X[2] - X[0], Y31 = Y[2] - Y[0], Z31 = Z[2] - Z[0];
  double SL21 = sqrt(X21 * X21 + Y21 * Y21 + Z21 * Z21);
  double SL31 = sqrt(X31 * X31 + Y31 * Y31 + Z31 * Z31);
  double COSG = (X31 * X21 + Y31 * Y21 + Z31 * Z21) / (SL31 * SL21);
  double AUX1 = 1. - COSG * COSG;
  double SING = AUX1 <= 0.0 ? 0. : sqrt(AUX1);
  XL[0] = 0.;  XL[1] = SL21;  XL[2] = SL31 * COSG;
  YL[0] = 0.;  YL[1] = 0.;    YL[2] = SL31 * SING;
}

/* hlst.f:7-293, only the kappa matrix AK(7,9) (column-major) is produced; the 9x9
 * stiffness EK is not on the stress path. */
static int hlst31(double *AK, const double *E, const double *X, const double *Y, double THK,
                  int IOP)
{
  static const int IP[3] = {2, 3, 1};
  double C[3], S[3], SL[3], EI[9], A[7 * 9], F7[49], AUX[9], XL[3], YL[3];
  double A1, A4, A5, AREA, B1, B2, B3, B4, B5, C2, CX, CY, F, P02, P11, P20, S2, SC, SX, SY,
      X0, Y0;
#define EIm(i, j) EI[((i)-1) + 3 * ((j)-1)]
#define Fm(i, j) F7[((i)-1) + 7 * ((j)-1)]
#define Am(i, j) A[((i)-1) + 7 * ((j)-1)]
#define AKm(i, j) AK[((i)-1) + 7 * ((j)-1)]
  for (int i = 0; i < 3; i++) {
    int j = IP[i] - 1;
    SL[i] = sqrt((X[j] - X[i]) * (X[j] - X[i]) + (Y[j] - Y[i]) * (Y[j] - Y[i]));
  }
  AREA = X[0] * Y[1] + X[1] * Y[2] + X[2] * Y[0];
  AREA = AREA - X[0] * Y[2] - X[1] * Y[0] - X[2] * Y[1];
  AREA = 0.5 * AREA;
  if (AREA < 0.0) return -1;

  for (int i = 0; i < 3; i++) {
    int j = IP[i] - 1;
    S[i] = (X[i] - X[j]) / SL[i];
    C[i] = (Y[j] - Y[i]) / SL[i];
  }
  X0 = (X[0] + X[1] + X[2]) / 3.;
  Y0 = (Y[0] + Y[1] + Y[2]) / 3.;
  for (int i = 0; i < 3; i++) { XL[i] = X[i] - X0; YL[i] = Y[i] - Y0; }
  F = AREA / (12. * THK);
  P20 = F * (XL[0] * XL[0] + XL[1] * XL[1] + XL[2] * XL[2]);
  P11 = F * (XL[0] * YL[0] + XL[1] * YL[1] + XL[2] * YL[2]);
  P02 = F * (YL[0] * YL[0] + YL[1] * YL[1] + YL[2] * YL[2]);

  memcpy(EI, E, sizeof(EI));
  if (dinv12(3, EI) != 0) return -3;

  memset(F7, 0, sizeof(F7));
  F = AREA / THK;
  Fm(1, 1) = F * EIm(1, 1);
  Fm(1, 2) = F * EIm(1, 2);
  Fm(2, 2) = F * EIm(2, 2);
  Fm(1, 3) = F * EIm(1, 3);
  Fm(2, 3) = F * EIm(2, 3);
  Fm(3, 3) = F * EIm(3, 3);
  Fm(4, 4) = EIm(1, 1) * P20 - 2. * EIm(1, 3) * P11 + EIm(3, 3) * P02;
  Fm(4, 5) = EIm(1, 2) * P20 - EIm(2, 3) * P11;
  Fm(5, 5) = EIm(2, 2) * P20;
  Fm(4, 6) = EIm(1, 1) * P11 - EIm(1, 3) * P02;
  Fm(5, 6) = EIm(1, 2) * P11;
  Fm(6, 6) = EIm(1, 1) * P02;
  Fm(4, 7) = -EIm(1, 3) * P20 + (EIm(1, 2) + EIm(3, 3)) * P11 - EIm(2, 3) * P02;
  Fm(5, 7) = EIm(2, 2) * P11 - EIm(2, 3) * P20;
  Fm(6, 7) = EIm(1, 2) * P02 - EIm(1, 3) * P11;
  Fm(7, 7) = EIm(2, 2) * P02 - 2. * EIm(2, 3) * P11 + EIm(3, 3) * P20;
  for (int i = 1; i <= 7; i++)
    for (int j = i; j <= 7; j++) Fm(j, i) = Fm(i, j);
  if (dinv12(7, F7) != 0) return -4;

  memset(A, 0, sizeof(A));
  for (int K = 1; K <= 3; K++) {
    double c = C[K - 1], s = S[K - 1], sl = SL[K - 1];
    C2 = c * c;  S2 = s * s;  SC = s * c;
    CX = c * XL[K - 1];  CY = c * YL[K - 1];  SX = s * XL[K - 1];  SY = s * YL[K - 1];
    F = sl * sl;
    A1 = sl / 2.;
    A4 = -c * F / 12.;
    A5 = -s * F / 12.;
    for (int I = 1; I <= 2; I++) {
      int J;
      if (I == 1) {
        J = 3 * K - 2;
        B4 = -c * F * sl / 30.;
        B5 = -s * F * sl / 30.;
        if (IOP <= 0) {
          B1 = F * (9. * C2 + 10. * S2) / 60.;
          B2 = -F * SC / 60.;
          B3 = F * (10. * C2 + 9. * S2) / 60.;
        } else {
          B1 = F / 6.;  B2 = 0.;  B3 = B1;
        }
      } else {
        J = 3 * IP[K - 1] - 2;
        A4 = -A4;
        A5 = -A5;
        B4 = c * F * sl / 20.;
        B5 = s * F * sl / 20.;
        if (IOP <= 0) {
          B1 = F * (21. * C2 + 20. * S2) / 60.;
          B2 = F * SC / 60.;
          B3 = F * (20. * C2 + 21. * S2) / 60.;
        } else {
          B1 = F / 3.;  B2 = 0.;  B3 = B1;
        }
      }
      Am(1, J) = Am(1, J) + A1 * c;
      Am(3, J) = Am(3, J) + A1 * s;
      Am(4, J) = Am(4, J) + A1 * (CX - SY) - 2. * B1 * SC - B2 * C2;
      Am(5, J) = Am(5, J) - B2 * S2;
      Am(6, J) = Am(6, J) + A1 * CY + B1 * C2;
      Am(7, J) = Am(7, J) - A1 * SX + B1 * S2 + 2. * B2 * SC;
      Am(2, J + 1) = Am(2, J + 1) + A1 * s;
      Am(3, J + 1) = Am(3, J + 1) + A1 * c;
      Am(4, J + 1) = Am(4, J + 1) - A1 * CY - 2. * B2 * SC - B3 * C2;
      Am(5, J + 1) = Am(5, J + 1) + A1 * SX - B3 * S2;
      Am(6, J + 1) = Am(6, J + 1) + B2 * C2;
      Am(7, J + 1) = Am(7, J + 1) + A1 * (SY - CX) + B2 * S2 + 2. * B3 * SC;
      Am(1, J + 2) = Am(1, J + 2) + A4 * c;
      Am(2, J + 2) = Am(2, J + 2) + A5 * s;
      Am(3, J + 2) = Am(3, J + 2) + A4 * s + A5 * c;
      Am(4, J + 2) = Am(4, J + 2) + A4 * (CX - SY) - A5 * CY - 2. * B4 * SC - B5 * C2;
      Am(5, J + 2) = Am(5, J + 2) + A5 * SX - B5 * S2;
      Am(6, J + 2) = Am(6, J + 2) + A4 * CY + B4 * C2;
      Am(7, J + 2) = Am(7, J + 2) - A4 * SX + A5 * (SY - CX) + B4 * S2 + 2. * B5 * SC;
    }
  }
  /* kappa = F^-1 * A */
  for (int I = 1; I <= 7; I++) {
    for (int J = 1; J <= 9; J++) {
      AUX[J - 1] = Fm(I, 1) * Am(1, J);
      for (int K = 2; K <= 7; K++) AUX[J - 1] = AUX[J - 1] + Fm(I, K) * Am(K, J);
    }
    for (int J = 1; J <= 9; J++) AKm(I, J) = AUX[J - 1];
  }
#undef EIm
#undef Fm
#undef Am
#undef AKm
  return 0;
}

/* hlst.f:294-355; SM(3,9) column-major */
static void hlst32(double *SM, const double *AK, const double *X, const double *Y,
                   const double *Z)
{
  double RN[3][7], XL[3], YL[3], X0, Y0;
  X0 = (X[0] + X[1] + X[2]) / 3.;
  Y0 = (Y[0] + Y[1] + Y[2]) / 3.;
  for (int i = 0; i < 3; i++) { XL[i] = X[i] - X0; YL[i] = Y[i] - Y0; }
  memset(RN, 0, sizeof(RN));
  for (int i = 0; i < 3; i++) {
    RN[0][3] = RN[0][3] + XL[i] * Z[i];
    RN[0][5] = RN[0][5] + YL[i] * Z[i];
    RN[1][4] = RN[1][4] + XL[i] * Z[i];
    RN[1][6] = RN[1][6] + YL[i] * Z[i];
    RN[2][3] = RN[2][3] - YL[i] * Z[i];
    RN[2][6] = RN[2][6] - XL[i] * Z[i];
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 9; j++) {
      double s = AK[i + 7 * j];
      for (int k = 3; k < 7; k++) s = s + RN[i][k] * AK[k + 7 * j];
      SM[i + 3 * j] = s;
    }
}

/* nyteba.f:7-331, only the kappa matrix AK(9,9) (column-major) is produced. */
static int teba31(double *AK, const double *E, const double *X, const double *Y,
                  const double *TH)
{
  static const int IP[3] = {2, 3, 1};
  /* nyteba.f:35-43: REAL*4 DATA constants */
  const double Z1[7] = {F32(0.33333333f), F32(0.05971587f), F32(0.47014206f),
                        F32(0.47014206f), F32(0.79742699f), F32(0.10128651f),
                        F32(0.10128651f)};
  const double Z2[7] = {F32(0.33333333f), F32(0.47014206f), F32(0.05971587f),
                        F32(0.47014206f), F32(0.10128651f), F32(0.79742699f),
                        F32(0.10128651f)};
  const double Z3[7] = {F32(0.33333333f), F32(0.47014206f), F32(0.47014206f),
                        F32(0.05971587f), F32(0.10128651f), F32(0.10128651f),
                        F32(0.79742699f)};
  const double W[7] = {F32(0.225f),      F32(0.13239415f), F32(0.13239415f), F32(0.13239415f),
                       F32(0.12593918f), F32(0.12593918f), F32(0.12593918f)};
  double C[3], S[3], SL[3], EI[9], B[9], G[9 * 12], T[12 * 9], EK[81], GT[81];
  double F, A, RL11, RL12, RL21, RL22, THK, SS, CC, CS;
#define EIm(i, j) EI[((i)-1) + 3 * ((j)-1)]
#define Bm(i, j) B[((i)-1) + 3 * ((j)-1)]
#define Gm(i, j) G[((i)-1) + 9 * ((j)-1)]
#define Tm(i, j) T[((i)-1) + 12 * ((j)-1)]
#define EKm(i, j) EK[((i)-1) + 9 * ((j)-1)]
  for (int i = 0; i < 3; i++) {
    int j = IP[i] - 1;
    SL[i] = sqrt((X[j] - X[i]) * (X[j] - X[i]) + (Y[j] - Y[i]) * (Y[j] - Y[i]));
  }
  A = X[0] * Y[1] + X[1] * Y[2] + X[2] * Y[0];
  A = A - X[0] * Y[2] - X[1] * Y[0] - X[2] * Y[1];
  A = 0.5 * A;
  if (A <= 0.0) return -1;
  for (int i = 0; i < 3; i++) {
    int j = IP[i] - 1;
    S[i] = (X[i] - X[j]) / SL[i];
    C[i] = (Y[j] - Y[i]) / SL[i];
  }
  RL11 = 0.5 * (Y[1] - Y[2]) / A;
  RL12 = 0.5 * (Y[2] - Y[0]) / A;
  RL21 = 0.5 * (X[2] - X[1]) / A;
  RL22 = 0.5 * (X[0] - X[2]) / A;

  memcpy(EI, E, sizeof(EI));
  if (dinv12(3, EI) != 0) return -3;

  memset(B, 0, sizeof(B));
  for (int K = 0; K < 7; K++) {
    THK = TH[0] * Z1[K] + TH[1] * Z2[K] + TH[2] * Z3[K];
    if (THK <= 0.0) return -2;
    F = 12. * A * W[K] / (THK * THK * THK);
    Bm(1, 1) = Bm(1, 1) + F * Z1[K] * Z1[K];
    Bm(1, 2) = Bm(1, 2) + F * Z1[K] * Z2[K];
    Bm(1, 3) = Bm(1, 3) + F * Z1[K] * Z3[K];
    Bm(2, 2) = Bm(2, 2) + F * Z2[K] * Z2[K];
    Bm(2, 3) = Bm(2, 3) + F * Z2[K] * Z3[K];
    Bm(3, 3) = Bm(3, 3) + F * Z3[K] * Z3[K];
  }
  for (int i = 1; i <= 3; i++)
    for (int j = i; j <= 3; j++) Bm(j, i) = Bm(i, j);

  for (int II = 1; II <= 3; II++)
    for (int JJ = II; JJ <= 3; JJ++) {
      int MM = 3 * II - 3, NN = 3 * JJ - 3;
      for (int i = 1; i <= 3; i++)
        for (int j = 1; j <= 3; j++) EKm(MM + i, NN + j) = EIm(II, JJ) * Bm(i, j);
    }
  for (int i = 1; i <= 9; i++)
    for (int j = i; j <= 9; j++) EKm(j, i) = EKm(i, j);
  if (dinv12(9, EK) != 0) return -4;

  memset(G, 0, sizeof(G));
  memset(T, 0, sizeof(T));
  for (int I = 1; I <= 3; I++) {
    int J = IP[I - 1], K = IP[J - 1];
    Gm(I, I) = S[K - 1] * C[K - 1] - S[I - 1] * C[I - 1];
    Gm(I + 3, I) = -Gm(I, I);
    Gm(I + 6, I) = C[I - 1] * C[I - 1] - S[I - 1] * S[I - 1] - C[K - 1] * C[K - 1] +
                   S[K - 1] * S[K - 1];
  }
  for (int I = 1; I <= 3; I++) {
    int J = I + 3;
    double c = C[I - 1], s = S[I - 1];
    SS = s * s * c;
    CC = c * c * s;
    CS = c * c - s * s;
    Gm(1, J) = (c + SS) * RL11 - CC * RL21;
    Gm(2, J) = (c + SS) * RL12 - CC * RL22;
    Gm(3, J) = -(c + SS) * (RL11 + RL12) + CC * (RL21 + RL22);
    Gm(4, J) = (s + CC) * RL21 - SS * RL11;
    Gm(5, J) = (s + CC) * RL22 - SS * RL12;
    Gm(6, J) = -(s + CC) * (RL21 + RL22) + SS * (RL11 + RL12);
    Gm(7, J) = s * RL11 + c * RL21 - CS * (s * RL11 - c * RL21);
    Gm(8, J) = c * RL22 + s * RL12 - CS * (s * RL12 - c * RL22);
    Gm(9, J) = -s * (RL11 + RL12) - c * (RL21 + RL22);
    Gm(9, J) = Gm(9, J) + CS * (s * (RL11 + RL12) - c * (RL21 + RL22));
  }
  for (int I = 1; I <= 3; I++) {
    int J = IP[I - 1], N = 2 * I + 5;
    double c = C[I - 1], s = S[I - 1];
    Gm(I, N) = c * c;
    Gm(I + 3, N) = s * s;
    Gm(I + 6, N) = 2. * s * c;
    Gm(J, N + 1) = c * c;
    Gm(J + 3, N + 1) = s * s;
    Gm(J + 6, N + 1) = 2. * s * c;
  }
  for (int I = 1; I <= 3; I++) {
    int J = IP[I - 1];
    double sl = SL[I - 1];
    SS = S[I - 1] * sl;
    CC = C[I - 1] * sl;
    Tm(I, 3 * I - 2) = 1.;
    Tm(I + 3, 3 * I - 2) = sl / 2.;
    Tm(I + 3, 3 * I - 1) = -sl * SS / 12.;
    Tm(I + 3, 3 * I) = sl * CC / 12.;
    Tm(I + 3, 3 * J - 2) = sl / 2.;
    Tm(I + 3, 3 * J - 1) = sl * SS / 12.;
    Tm(I + 3, 3 * J) = -sl * CC / 12.;
    Tm(2 * I + 5, 3 * I - 1) = -CC / 3.;
    Tm(2 * I + 5, 3 * I) = -SS / 3.;
    Tm(2 * I + 5, 3 * J - 1) = -CC / 6.;
    Tm(2 * I + 5, 3 * J) = -SS / 6.;
    Tm(2 * I + 6, 3 * I - 1) = -CC / 6.;
    Tm(2 * I + 6, 3 * I) = -SS / 6.;
    Tm(2 * I + 6, 3 * J - 1) = -CC / 3.;
    Tm(2 * I + 6, 3 * J) = -SS / 3.;
  }
  for (int I = 1; I <= 3; I++) {
    int J = 3 * I - 1, K = J + 1;
    for (int M = 1; M <= 12; M++) {
      F = Tm(M, J);
      Tm(M, J) = Tm(M, K);
      Tm(M, K) = -F;
    }
  }
  /* A = G*T (9x9), then kappa = EK^-1-flex * A */
  for (int I = 1; I <= 9; I++)
    for (int J = 1; J <= 9; J++) {
      double s = 0.;
      for (int K = 1; K <= 12; K++) s = s + Gm(I, K) * Tm(K, J);
      GT[(I - 1) + 9 * (J - 1)] = s;
    }
  for (int I = 1; I <= 9; I++)
    for (int J = 1; J <= 9; J++) {
      double s = 0.;
      for (int K = 1; K <= 9; K++) s = s + EKm(I, K) * GT[(K - 1) + 9 * (J - 1)];
      AK[(I - 1) + 9 * (J - 1)] = s;
    }
#undef EIm
#undef Bm
#undef Gm
#undef Tm
#undef EKm
  return 0;
}

/* nyteba.f:333-364 */
static void teba32(double *SM, const double *AK, const double *Z)
{
  for (int I = 1; I <= 3; I++) {
    int L = 3 * I - 3;
    for (int J = 1; J <= 9; J++) {
      double s = 0.;
      for (int K = 1; K <= 3; K++) s = s + Z[K - 1] * AK[(L + K - 1) + 9 * (J - 1)];
      SM[(I - 1) + 3 * (J - 1)] = s;
    }
  }
}

/* fts.f:446-490 */
static void fts38(double *VML, double *VBL, const double *V, const double *X, const double *Y,
                  const double *Z)
{
  double A[3], B[3], C[9];
  dirc30(C, X, Y, Z);
  for (int I = 1; I <= 3; I++) {
    A[0] = V[6 * I - 6]; A[1] = V[6 * I - 5]; A[2] = V[6 * I - 4];
    for (int L = 0; L < 3; L++) {
      B[L] = C[L + 0] * A[0];
      for (int K = 1; K < 3; K++) B[L] = B[L] + C[L + 3 * K] * A[K];
    }
    VML[3 * I - 3] = B[0];
    VML[3 * I - 2] = B[1];
    VBL[3 * I - 3] = B[2];
    A[0] = V[6 * I - 3]; A[1] = V[6 * I - 2]; A[2] = V[6 * I - 1];
    for (int L = 0; L < 3; L++) {
      B[L] = C[L + 0] * A[0];
      for (int K = 1; K < 3; K++) B[L] = B[L] + C[L + 3 * K] * A[K];
    }
    VBL[3 * I - 2] = B[0];
    VBL[3 * I - 1] = B[1];
    VML[3 * I - 1] = B[2];
  }
}

/* tmrf.f:304-388 (LUFACT): Crout factorisation with implicit row scaling, A(NA,*) column-major, diagonal stored inverted */
static int lufact9(double *A, int *IPERM, double *V)
{
  const int N = 9;
  const double MACTOL = 2.0e-16;
#define A_(i, j) A[((i) - 1) + 9 * ((j) - 1)]
  for (int I = 1; I <= N; I++) {
    double Y = 0.0;
    for (int J = 1; J <= N; J++) Y = Y + A_(I, J) * A_(I, J);
    V[I - 1] = Y > 0.0 ? sqrt(1.0 / Y) : 0.0;
  }
  for (int K = 1; K <= N; K++) {
    IPERM[K - 1] = K;
    if (V[K - 1] <= 0.0) continue;
    int L = K;
    double X = 0.0;
    for (int I = K; I <= N; I++) {
      double Y = 0.0;
      for (int J = 1; J <= K - 1; J++) Y = Y + A_(I, J) * A_(J, K);
      A_(I, K) = A_(I, K) - Y;
      Y = fabs(V[I - 1] * A_(I, K));
      if (Y > X) { X = Y; L = I; }
    }
    if (L != K) {
      for (int J = 1; J <= N; J++) { double Y = A_(K, J); A_(K, J) = A_(L, J); A_(L, J) = Y; }
      V[L - 1] = V[K - 1];
      IPERM[K - 1] = L;
    }
    if (X <= MACTOL) return N - (K - 1);
    X = 1.0 / A_(K, K);
    A_(K, K) = X;
    for (int J = K + 1; J <= N; J++) {
      double Y = 0.0;
      for (int I = 1; I <= K - 1; I++) Y = Y + A_(K, I) * A_(I, J);
      A_(K, J) = (A_(K, J) - Y) * X;
    }
  }
  return 0;
}

/* tmrf.f:389-473 (LUSOLV) for the call LUSOLV(GT,9,9,IPERM,HH,3,-3): three right hand sides stored as ROWS of B(3,9) */
static void lusolv9_rows(const double *A, const int *IPERM, double *B)
{
  const int N = 9;
#define B_(j, i) B[((j) - 1) + 3 * ((i) - 1)]
  for (int I = 1; I <= N; I++) {
    const int K = IPERM[I - 1];
    if (K != I)
      for (int J = 1; J <= 3; J++) { double T = B_(J, I); B_(J, I) = B_(J, K); B_(J, K) = T; }
  }
  for (int J = 1; J <= 3; J++) {
    B_(J, 1) = B_(J, 1) * A_(1, 1);
    for (int I = 1; I <= N - 1; I++) {
      double SUM = 0.0;
      for (int K = 1; K <= I; K++) SUM = SUM - A_(I + 1, K) * B_(J, K);
      B_(J, I + 1) = (B_(J, I + 1) + SUM) * A_(I + 1, I + 1);
    }
    for (int I = N - 1; I >= 1; I--) {
      double SUM = 0.0;
      for (int K = 1; K <= N - I; K++) SUM = SUM - A_(I, I + K) * B_(J, I + K);
      B_(J, I) = B_(J, I) + SUM;
    }
  }
#undef B_
#undef A_
}

/* The HH(3,9) output of TMRF31 / SM3MH (tmrf.f:7-55,151-303) as FTS31 calls it (IAT = 0: one triangle, BETA > 0): the
 * higher-order strain-displacement relation of the Bergan / Felippa membrane triangle; columns in SM3MH's own DOF order
 * (u1 v1 u2 v2 u3 v3 th1 th2 th3).  The stiffness part of the routine does not feed the stresses and is left out. */
static int tmrf31_hh(double *HH, const double *X, const double *Y)
{
  double GT[81], T[9], XC[3], YC[3], XM[3], YM[3];
  int IPERM[9];
  const double AREA2 = (Y[1] - Y[0]) * (X[0] - X[2]) - (X[1] - X[0]) * (Y[0] - Y[2]);
  if (AREA2 <= 1.0e-16) return -1; /* NEGA_AREA / ZERO_AREA */
  const double X0 = (X[0] + X[1] + X[2]) / 3.0, Y0 = (Y[0] + Y[1] + Y[2]) / 3.0;
  const double AREA = 0.5 * AREA2;
  const double C = 1. / sqrt(AREA);
  for (int i = 0; i < 3; i++) { XC[i] = C * (X[i] - X0); YC[i] = C * (Y[i] - Y0); }
  XM[0] = 0.5 * (XC[1] + XC[2]); XM[1] = 0.5 * (XC[2] + XC[0]); XM[2] = 0.5 * (XC[0] + XC[1]);
  YM[0] = 0.5 * (YC[1] + YC[2]); YM[1] = 0.5 * (YC[2] + YC[0]); YM[2] = 0.5 * (YC[0] + YC[1]);
#define GT_(i, j) GT[((i) - 1) + 9 * ((j) - 1)]
#define HH_(i, j) HH[((i) - 1) + 3 * ((j) - 1)]
  /* only rows 1..6 of GT are zeroed (tmrf.f:206-212); rows 7..9 are fully assigned below */
  for (int I = 1; I <= 9; I++) {
    for (int J = 1; J <= 6; J++) GT_(J, I) = 0.;
    HH_(1, I) = 0.; HH_(2, I) = 0.; HH_(3, I) = 0.;
  }
  for (int J = 1; J <= 3; J++) {
    const double DX = XM[J - 1] - XC[J - 1], DY = YM[J - 1] - YC[J - 1];
    const double DL = sqrt(DX * DX + DY * DY);
    const double CJ = DX / DL, SJ = DY / DL;
    const double A1J = -0.5 * SJ * (CJ * CJ), A2J = 0.5 * (CJ * CJ * CJ), B2J = -0.5 * (SJ * SJ * SJ), B3J = 0.5 * (SJ * SJ) * CJ;
    const double A3J = -(B2J + A1J + A1J), B1J = -(B3J + B3J + A2J);
    GT_(1, 2 * J - 1) = 1.;
    GT_(2, 2 * J) = 1.;
    GT_(3, 2 * J - 1) = -YC[J - 1];
    GT_(3, 2 * J) = XC[J - 1];
    GT_(3, J + 6) = C;
    GT_(4, 2 * J - 1) = XC[J - 1];
    GT_(6, 2 * J - 1) = YC[J - 1];
    GT_(5, 2 * J) = YC[J - 1];
    GT_(6, 2 * J) = XC[J - 1];
    HH_(J, J + 6) = 1.;
    for (int I = 1; I <= 3; I++) {
      const double XI = XC[I - 1], YI = YC[I - 1];
      GT_(J + 6, 2 * I - 1) = A1J * XI * XI + 2. * A2J * XI * YI + A3J * YI * YI;
      GT_(J + 6, 2 * I) = B1J * XI * XI + 2. * B2J * XI * YI + B3J * YI * YI;
      GT_(J + 6, I + 6) = -C * (CJ * XI + SJ * YI);
    }
  }
  if (lufact9(GT, IPERM, T) != 0) return -2; /* SINGULAR_G */
  lusolv9_rows(GT, IPERM, HH);
#undef GT_
#undef HH_
  return 0;
}

/* tmrf.f:474-634 (TMRF32): the centroid stress matrix SM(3,9) = DM * (L'/AREA + BH * HH).  L and TMB run over the DOFs in node
 * order (u1 v1 th1 u2 ..) while the columns of HH come in SM3MH's order (u1 v1 u2 v2 u3 v3 th1 th2 th3): the reference adds the
 * two column by column as they are, and so does this restatement. */
static int tmrf32(const double *HH, const double *X, const double *Y, const double *DM, double ALPHA, double *SM)
{
  double L[27], TMB[27], TMH[27], BH[9], TM[27], XC[3], YC[3], XM[3], YM[3];
  const double X21 = X[1] - X[0], X12 = -X21, X32 = X[2] - X[1], X23 = -X32, X13 = X[0] - X[2], X31 = -X13;
  const double Y21 = Y[1] - Y[0], Y12 = -Y21, Y32 = Y[2] - Y[1], Y23 = -Y32, Y13 = Y[0] - Y[2], Y31 = -Y13;
  const double AREA = 0.5 * (Y21 * X13 - X21 * Y13);
  if (AREA <= 1.0e-16) return -1;
#define L_(i, j) L[((i) - 1) + 9 * ((j) - 1)]
  for (int i = 0; i < 27; i++) L[i] = 0.0; /* rows 3, 6, 9 are only set for ALPHA > 0 (always, here) */
  L_(1, 1) = 0.5 * Y23; L_(2, 1) = 0.5 * 0.0; L_(4, 1) = 0.5 * Y31; L_(5, 1) = 0.5 * 0.0; L_(7, 1) = 0.5 * Y12; L_(8, 1) = 0.5 * 0.0;
  L_(1, 2) = 0.5 * 0.0; L_(2, 2) = 0.5 * X32; L_(4, 2) = 0.5 * 0.0; L_(5, 2) = 0.5 * X13; L_(7, 2) = 0.5 * 0.0; L_(8, 2) = 0.5 * X21;
  L_(1, 3) = 0.5 * X32; L_(2, 3) = 0.5 * Y23; L_(4, 3) = 0.5 * X13; L_(5, 3) = 0.5 * Y31; L_(7, 3) = 0.5 * X21; L_(8, 3) = 0.5 * Y12;
  if (ALPHA > 0.0) {
    L_(3, 1) = 0.5 * (Y23 * (Y13 - Y21) * ALPHA / 6);
    L_(3, 2) = 0.5 * (X32 * (X31 - X12) * ALPHA / 6);
    L_(3, 3) = 0.5 * ((X31 * Y13 - X12 * Y21) * ALPHA / 3);
    L_(6, 1) = 0.5 * (Y31 * (Y21 - Y32) * ALPHA / 6);
    L_(6, 2) = 0.5 * (X13 * (X12 - X23) * ALPHA / 6);
    L_(6, 3) = 0.5 * ((X12 * Y21 - X23 * Y32) * ALPHA / 3);
    L_(9, 1) = 0.5 * (Y12 * (Y32 - Y13) * ALPHA / 6);
    L_(9, 2) = 0.5 * (X21 * (X23 - X31) * ALPHA / 6);
    L_(9, 3) = 0.5 * ((X23 * Y32 - X31 * Y13) * ALPHA / 3);
  }
  for (int I = 1; I <= 9; I++)
    for (int J = 1; J <= 3; J++) TMB[(J - 1) + 3 * (I - 1)] = (1. / AREA) * L_(I, J);
#undef L_
  const double X0 = (X[0] + X[1] + X[2]) / 3.0, Y0 = (Y[0] + Y[1] + Y[2]) / 3.0;
  const double C = 1. / sqrt(AREA);
  for (int i = 0; i < 3; i++) { XC[i] = C * (X[i] - X0); YC[i] = C * (Y[i] - Y0); }
  XM[0] = 0.5 * (XC[1] + XC[2]); XM[1] = 0.5 * (XC[2] + XC[0]); XM[2] = 0.5 * (XC[0] + XC[1]);
  YM[0] = 0.5 * (YC[1] + YC[2]); YM[1] = 0.5 * (YC[2] + YC[0]); YM[2] = 0.5 * (YC[0] + YC[1]);
  for (int J = 1; J <= 3; J++) {
    const double DX = XM[J - 1] - XC[J - 1], DY = YM[J - 1] - YC[J - 1];
    const double DL = sqrt(DX * DX + DY * DY);
    const double CJ = DX / DL, SJ = DY / DL;
    const double A1J = -0.5 * SJ * (CJ * CJ), A2J = 0.5 * (CJ * CJ * CJ), B2J = -0.5 * (SJ * SJ * SJ), B3J = 0.5 * (SJ * SJ) * CJ;
    BH[0 + 3 * (J - 1)] = C * (2 * A1J * XC[J - 1] + A2J * YC[J - 1]);
    BH[1 + 3 * (J - 1)] = C * (B2J * XC[J - 1] + 2 * B3J * YC[J - 1]);
    BH[2 + 3 * (J - 1)] = C * (-4 * B3J * XC[J - 1] - 4 * A1J * YC[J - 1]);
  }
  for (int I = 0; I < 9; I++)
    for (int J = 0; J < 3; J++) {
      double S = 0.0;
      for (int K = 0; K < 3; K++) S = S + BH[J + 3 * K] * HH[K + 3 * I];
      TMH[J + 3 * I] = S;
    }
  for (int i = 0; i < 27; i++) TM[i] = TMB[i] + TMH[i];
  for (int I = 0; I < 9; I++)
    for (int J = 0; J < 3; J++) {
      double S = 0.0;
      for (int K = 0; K < 3; K++) S = S + DM[J + 3 * K] * TM[K + 3 * I];
      SM[J + 3 * I] = S;
    }
  return 0;
}

/* TEST HOOK: the centroid membrane stress matrix of the TMRF triangle, SM(3,9) over (u1 v1 th1 u2 v2 th2 u3 v3 th3), from the two
 * routines above.  reorder = 0: exactly what FTS32 computes (HH columns taken as they come); reorder = 1: HH columns first moved to
 * the node order TMRF32's lumping matrix uses -- the consistent element, which must reproduce every constant strain state.  The
 * closed-form test on the latter checks the restatement of HH and BH without a compiled reference. */
int orc_tmrf_centroid_matrix(const double XL[3], const double YL[3], const double DM[9], double alpha, int reorder, double SM[27])
{
  double HH[27], H2[27];
  if (tmrf31_hh(HH, XL, YL) < 0) return -1;
  if (reorder) {
    static const int LST[9] = {1, 2, 4, 5, 7, 8, 3, 6, 9};
    for (int j = 0; j < 9; j++)
      for (int r = 0; r < 3; r++) H2[r + 3 * (LST[j] - 1)] = HH[r + 3 * j];
    memcpy(HH, H2, sizeof(HH));
  }
  return tmrf32(HH, XL, YL, DM, alpha, SM);
}

/* -fftStressForm of the legacy FFT shell (type 21): 1 = FTSA31 / FTSA32 (default, the statements of STR23), anything else =
 * FTS31 / FTS32 (elStressModule.f90:559-566,586-592) */
static int g_fft_stress_form = 1;
void orc_set_fft_stress_form(int form) { g_fft_stress_form = form; }
int orc_get_fft_stress_form(void) { return g_fft_stress_form; }

/* STR21 with -fftStressForm /= 1 (elStressModule.f90:521-633): FTS31 (fts.f:7-213) delivers HH (membrane, TMRF31 with ALPHA = 1.5,
 * BETA = 0.5) and the kappa matrix AKB of TEBA31; FTS32 (fts.f:214-292) the centroid stress matrices, called with the plane stress
 * matrix E where it expects the membrane rigidity t * E -- kept as the reference has it; ZZ = 1/3 in double precision here. */
int orc_str21_legacy(const double xg[3], const double yg[3], const double zg[3], double emod,
                     double rny, const double thk[3], const double ev[18], double SR[18],
                     double SS[18], double sigma[18], double epsil[18])
{
  double E[9], HH[27], AKB[81], ESMM[27], ESMB[27], VML[9], VBL[9], RMF[3], RBF[3];
  double VX[3], VY[3], VZ[3], T_str[4], XL[3], YL[3], ZZ[3];
  const double ALPHA = 1.5;
  int ierr;

  orc_iso_mat2d(emod, rny, E);
  local_xy(xg, yg, zg, XL, YL);
  if (tmrf31_hh(HH, XL, YL) < 0) return 1;
  if (teba31(AKB, E, XL, YL, thk) < 0) return 1;

  ierr = orc_shell_element_axes(3, xg, yg, zg, VX, VY, VZ);
  if (ierr != 0) return ierr;
  ierr = orc_shell_stress_trans(VX, VZ, T_str);
  if (ierr != 0) return ierr;

  for (int i = 0; i < 3; i++) ZZ[i] = 1.0 / 3.0;
  local_xy(xg, yg, zg, XL, YL);
  if (tmrf32(HH, XL, YL, E, ALPHA, ESMM) < 0) return 1;
  teba32(ESMB, AKB, ZZ);

  fts38(VML, VBL, ev, xg, yg, zg);
  for (int i = 0; i < 3; i++) {
    double sm = 0.0, sb = 0.0;
    for (int j = 0; j < 9; j++) {
      sm += ESMM[i + 3 * j] * VML[j];
      sb += ESMB[i + 3 * j] * VBL[j];
    }
    RMF[i] = sm;
    RBF[i] = sb;
  }
  orc_rotate2d(RMF, T_str, RMF);
  orc_rotate2d(RBF, T_str, RBF);
  for (int n = 0; n < 3; n++)
    for (int c = 0; c < 3; c++) {
      SR[c + 6 * n] = RMF[c];
      SR[3 + c + 6 * n] = RBF[c];
      SS[c + 6 * n] = 0.0;
      SS[3 + c + 6 * n] = 0.0;
    }
  for (int i = 0; i < 3; i++)
    for (int c = 0; c < 3; c++) {
      sigma[c + 3 * i] = (RMF[c] + RBF[c] * 6.0 / thk[i]) / thk[i];
      sigma[c + 3 * (3 + i)] = (RMF[c] - RBF[c] * 6.0 / thk[i]) / thk[i];
    }
  orc_iso_mat2d_inv(emod, rny, E);
  for (int p = 0; p < 6; p++)
    for (int r = 0; r < 3; r++)
      epsil[r + 3 * p] = E[r] * sigma[3 * p] + E[r + 3] * sigma[1 + 3 * p] +
                         E[r + 6] * sigma[2 + 3 * p];
  return 0;
}

/* elStressModule.f90:901-999.  SR(6,3), sigma(3,6), epsil(3,6) column-major. */
int orc_str23(const double xg[3], const double yg[3], const double zg[3], double emod,
              double rny, const double thk[3], const double ev[18], double SR[18],
              double SS[18], double sigma[18], double epsil[18])
{
  double E[9], AKM[63], AKB[81], ESMM[27], ESMB[27], VML[9], VBL[9], RMF[3], RBF[3];
  double VX[3], VY[3], VZ[3], T_str[4], XL[3], YL[3], ZZ[3], THK;
  int ierr;

  orc_iso_mat2d(emod, rny, E);

  /* FTSA31 (ftsa.f:53-104): average thickness, local coordinates, HLST31 + TEBA31 */
  THK = (thk[0] + thk[1] + thk[2]) / 3.;
  local_xy(xg, yg, zg, XL, YL);
  if (hlst31(AKM, E, XL, YL, THK, 1) < 0) return 1;
  if (teba31(AKB, E, XL, YL, thk) < 0) return 1;

  ierr = orc_shell_element_axes(3, xg, yg, zg, VX, VY, VZ);
  if (ierr != 0) return ierr;
  ierr = orc_shell_stress_trans(VX, VZ, T_str);
  if (ierr != 0) return ierr;

  /* FTSA32 (ftsa.f:196-260) */
  for (int i = 0; i < 3; i++) ZZ[i] = (double)(1.f / 3.f); /* REAL*4 1./3., ftsa.f:223 */
  local_xy(xg, yg, zg, XL, YL);
  hlst32(ESMM, AKM, XL, YL, ZZ);
  teba32(ESMB, AKB, ZZ);

  fts38(VML, VBL, ev, xg, yg, zg);
  for (int i = 0; i < 3; i++) {
    double sm = 0.0, sb = 0.0;
    for (int j = 0; j < 9; j++) {
      sm += ESMM[i + 3 * j] * VML[j];
      sb += ESMB[i + 3 * j] * VBL[j];
    }
    RMF[i] = sm;
    RBF[i] = sb;
  }
  orc_rotate2d(RMF, T_str, RMF);
  orc_rotate2d(RBF, T_str, RBF);

  for (int n = 0; n < 3; n++)
    for (int c = 0; c < 3; c++) {
      SR[c + 6 * n] = RMF[c];
      SR[3 + c + 6 * n] = RBF[c];
      SS[c + 6 * n] = 0.0;
      SS[3 + c + 6 * n] = 0.0;
    }
  for (int i = 0; i < 3; i++)
    for (int c = 0; c < 3; c++) {
      sigma[c + 3 * i] = (RMF[c] + RBF[c] * 6.0 / thk[i]) / thk[i];
      sigma[c + 3 * (3 + i)] = (RMF[c] - RBF[c] * 6.0 / thk[i]) / thk[i];
    }
  orc_iso_mat2d_inv(emod, rny, E);
  for (int p = 0; p < 6; p++)
    for (int r = 0; r < 3; r++)
      epsil[r + 3 * p] = E[r] * sigma[3 * p] + E[r + 3] * sigma[1 + 3 * p] +
                         E[r + 6] * sigma[2 + 3 * p];
  return 0;
}
