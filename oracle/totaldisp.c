/*
 * totaldisp.c -- CPU restatement (TEST INFRASTRUCTURE, see oracle.h) of the total nodal displacement that
 * fedem_stress writes next to the deformational displacement when -deformation is on:
 *   calcTotalNodalDisplacement  src/vpmStress/displacementModule.f90:1694-1745
 *   vec_to_mat / mat_to_vec / vec_to_quat / quat_to_mat / mat_to_quat / quat_to_vec / deltaRot
 *                               src/vpmUtilities/rotationModule.f90:355-543
 *   matmul34                    src/vpmUtilities/manipMatrixModule.f90:311-322
 * Written with 1-based Fortran-style index macros so that it can be read side by side with the Fortran.
 * No reference golden exists for this routine (Fortran, cannot be built here): parity unpinned, checked
 * against an independent scipy.spatial rotation composition in tests/test_rdb_cpu.py.
 */
#include <math.h>
#include "oracle.h"

#define M3(a, i, j) a[(i - 1) + 3 * (j - 1)] /* column-major (3,n) */
static const double epsTh2_p = 0.0005;

static void vec_to_quat(const double *rvec, double *q)
{
  double thh = 0.5 * sqrt(rvec[0] * rvec[0] + rvec[1] * rvec[1] + rvec[2] * rvec[2]);
  double sthh = sin(thh), cthh = cos(thh), fac, f1, n;
  int i;
  if (thh > 1.0e6) {
    if (fabs(1.0 - cthh * cthh - sthh * sthh) > 0.00001) { q[0] = 1.0; q[1] = q[2] = q[3] = 0.0; return; }
  }
  if (thh < epsTh2_p) { f1 = thh / epsTh2_p; fac = f1 * sin(epsTh2_p) / epsTh2_p + 1.0 - f1; }
  else fac = sthh / thh;
  q[0] = cthh;
  for (i = 0; i < 3; ++i) q[1 + i] = rvec[i] * fac * 0.5;
  n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (i = 0; i < 4; ++i) q[i] /= n;
}

static void quat_to_mat(double *q, double *rten)
{
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  int i;
  for (i = 0; i < 4; ++i) q[i] /= n;
#define Q(i) q[i - 1]
  M3(rten, 1, 1) = 2.0 * (Q(2) * Q(2) + Q(1) * Q(1)) - 1.0;
  M3(rten, 2, 2) = 2.0 * (Q(3) * Q(3) + Q(1) * Q(1)) - 1.0;
  M3(rten, 3, 3) = 2.0 * (Q(4) * Q(4) + Q(1) * Q(1)) - 1.0;
  M3(rten, 1, 2) = 2.0 * (Q(2) * Q(3) - Q(4) * Q(1));
  M3(rten, 1, 3) = 2.0 * (Q(2) * Q(4) + Q(3) * Q(1));
  M3(rten, 2, 3) = 2.0 * (Q(3) * Q(4) - Q(2) * Q(1));
  M3(rten, 2, 1) = 2.0 * (Q(3) * Q(2) + Q(4) * Q(1));
  M3(rten, 3, 1) = 2.0 * (Q(4) * Q(2) - Q(3) * Q(1));
  M3(rten, 3, 2) = 2.0 * (Q(4) * Q(3) + Q(2) * Q(1));
}

static void mat_to_quat(const double *rten, double *q)
{
  double trace = M3(rten, 1, 1) + M3(rten, 2, 2) + M3(rten, 3, 3);
  int imax = 1, i, j, k;
  if (M3(rten, 2, 2) > M3(rten, imax, imax)) imax = 2;
  if (M3(rten, 3, 3) > M3(rten, imax, imax)) imax = 3;
  if (trace > M3(rten, imax, imax)) {
    Q(1) = sqrt(1.0 + trace) * 0.5;
    Q(2) = (M3(rten, 3, 2) - M3(rten, 2, 3)) / (4.0 * Q(1));
    Q(3) = (M3(rten, 1, 3) - M3(rten, 3, 1)) / (4.0 * Q(1));
    Q(4) = (M3(rten, 2, 1) - M3(rten, 1, 2)) / (4.0 * Q(1));
  } else {
    i = imax; j = imax % 3 + 1; k = (imax + 1) % 3 + 1;
    Q(i + 1) = sqrt(M3(rten, i, i) * 0.5 + (1.0 - trace) * 0.25);
    Q(1) = (M3(rten, k, j) - M3(rten, j, k)) / (4.0 * Q(i + 1));
    Q(j + 1) = (M3(rten, j, i) + M3(rten, i, j)) / (4.0 * Q(i + 1));
    Q(k + 1) = (M3(rten, k, i) + M3(rten, i, k)) / (4.0 * Q(i + 1));
  }
}

static void quat_to_vec(double *q, double *rvec)
{
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), cthh, sthh, thh, fac, f1;
  int i;
  for (i = 0; i < 4; ++i) q[i] /= n;
  cthh = Q(1);
  sthh = sqrt(Q(2) * Q(2) + Q(3) * Q(3) + Q(4) * Q(4));
  if (sthh < 0.7) thh = asin(sthh); else thh = acos(cthh);
  if (thh < epsTh2_p) { f1 = thh / epsTh2_p; fac = f1 * epsTh2_p / sin(epsTh2_p) + 1.0 - f1; }
  else if (sthh >= 1.0) fac = thh;
  else fac = thh / sthh;
  for (i = 0; i < 3; ++i) rvec[i] = q[1 + i] * fac * 2.0;
}

/* supTr, supTrInit: (3,4) column-major; uLoc/uTot: nd = 3 or 6 values */
void orc_total_nodal_displacement(const double *X0, const double *uLoc, int nd, const double *supTr,
                                  const double *supTrInit, double *uTot)
{
  double Xd[3], q[4], dR[9], A[9], B[9];
  int i, j, k;
  for (i = 1; i <= 3; ++i) Xd[i - 1] = X0[i - 1] + uLoc[i - 1];
  for (i = 1; i <= 3; ++i) {
    double xn = M3(supTr, i, 4), x0 = M3(supTrInit, i, 4);
    for (j = 1; j <= 3; ++j) { xn += M3(supTr, i, j) * Xd[j - 1]; x0 += M3(supTrInit, i, j) * X0[j - 1]; }
    uTot[i - 1] = xn - x0;
  }
  if (nd < 6) return;
  vec_to_quat(uLoc + 3, q);
  quat_to_mat(q, dR);
  for (i = 1; i <= 3; ++i)      /* matmul(dR, supTr): only its first three columns are used by deltaRot */
    for (j = 1; j <= 3; ++j) {
      double s = 0.0;
      for (k = 1; k <= 3; ++k) s += M3(dR, i, k) * M3(supTr, k, j);
      M3(A, i, j) = s;
    }
  for (i = 1; i <= 3; ++i)      /* deltaRot(T1,T2) = mat_to_vec(matmul(T2, transpose(T1))), T1 = supTrInit */
    for (j = 1; j <= 3; ++j) {
      double s = 0.0;
      for (k = 1; k <= 3; ++k) s += M3(A, i, k) * M3(supTrInit, j, k);
      M3(B, i, j) = s;
    }
  mat_to_quat(B, q);
  quat_to_vec(q, uTot + 3);
}
