/* rosette.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): strain rosettes / strain gages.
 * Follows src/vpmStress/strainRosetteModule.f90:506-580 (calcElmCoordSystem, explicit position
 * matrix as in the .fsi input format), :587-812 (InitStrainRosette: Teps, B_el, transformation to
 * global DOF directions, Bcart = bscr . H_el), :225-324 (evaluateStrainGages, calcRosetteStrains),
 * :327-353 (zero-start strains); src/vpmStress/strainGageModule.f90:604-661 (InitStrainGages: gage
 * direction vectors), :184-237 (rosette types); src/vpmStress/displacementModule.f90:1096-1202
 * (ElDispFromSupElDisp); src/vpmUtilities/rotationModule.f90:355-366,393-428,478-497
 * (vec_to_mat through quaternions); strainAndStressUtils.f90:14-98 (Mohr circle, in invariants.c). */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* rotationModule.f90:393-428 + 478-497 */
static void vec_to_mat(const double rvec[3], double rten[9] /* column-major 3x3 */)
{
  const double epsTh2 = 0.0005;
  double q[4], thh, sthh, cthh, f1, fac, nq;
  thh = 0.5 * sqrt(rvec[0] * rvec[0] + rvec[1] * rvec[1] + rvec[2] * rvec[2]);
  sthh = sin(thh);
  cthh = cos(thh);
  if (thh < epsTh2) {
    f1 = thh / epsTh2;
    fac = f1 * sin(epsTh2) / epsTh2 + 1.0 - f1;
  } else
    fac = sthh / thh;
  q[0] = cthh;
  for (int k = 0; k < 3; k++) q[k + 1] = rvec[k] * fac * 0.5;
  nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] = q[k] / nq;
  /* quat_to_mat normalises once more */
  nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] = q[k] / nq;
#define R(i, j) rten[((i)-1) + 3 * ((j)-1)]
  R(1, 1) = 2.0 * (q[1] * q[1] + q[0] * q[0]) - 1.0;
  R(2, 2) = 2.0 * (q[2] * q[2] + q[0] * q[0]) - 1.0;
  R(3, 3) = 2.0 * (q[3] * q[3] + q[0] * q[0]) - 1.0;
  R(1, 2) = 2.0 * (q[1] * q[2] - q[3] * q[0]);
  R(1, 3) = 2.0 * (q[1] * q[3] + q[2] * q[0]);
  R(2, 3) = 2.0 * (q[2] * q[3] - q[1] * q[0]);
  R(2, 1) = 2.0 * (q[2] * q[1] + q[3] * q[0]);
  R(3, 1) = 2.0 * (q[3] * q[1] - q[2] * q[0]);
  R(3, 2) = 2.0 * (q[3] * q[2] + q[1] * q[0]);
#undef R
}

/* strainGageModule.f90:645-661: Teps_NfromC of gage i = (c^2, s^2, c s).  Tg is [ngage][3]. */
void orc_gage_directions(const orc_rosette *ros, double *Tg)
{
  const double *X = ros->rpos, *Y = ros->rpos + 3, *Z = ros->rpos + 6;
  for (int i = 1; i <= ros->ngage; i++) {
    double rotVec[3], rotMat[9], v[3], c, s;
    for (int k = 0; k < 3; k++) rotVec[k] = (i - 1) * ros->alpha_gages * Z[k];
    vec_to_mat(rotVec, rotMat);
    for (int k = 0; k < 3; k++) v[k] = rotMat[k] * X[0] + rotMat[k + 3] * X[1] + rotMat[k + 6] * X[2];
    c = v[0] * X[0] + v[1] * X[1] + v[2] * X[2];
    s = v[0] * Y[0] + v[1] * Y[1] + v[2] * Y[2];
    Tg[3 * (i - 1)] = c * c;
    Tg[3 * (i - 1) + 1] = s * s;
    Tg[3 * (i - 1) + 2] = c * s;
  }
}

/* InitStrainRosette up to bscr (strainRosetteModule.f90:630-724).
 * bscr: 3 x nElDof column-major (at most 3 x 24); rows: the nodal DOF (1-based) behind each
 * column.  Returns 0, or the reference's error code. */
int orc_rosette_bscr(const orc_rosette *ros, const orc_sam *sam, const double *xyz, double *bscr,
                     int *rows, int *nElDof_out)
{
  const int nElNodes = ros->numnod;
  double X[4], Y[4], Z[4], T_el[9], V1[3], V2[3], V3[3], c[4], Teps[9], B_el[3 * 6 * 4];
  int nNDof = 0, nElDof = 0, ierr;
  if (nElNodes != 3 && nElNodes != 4) return -1;
  for (int i = 0; i < nElNodes; i++) {
    int node = ros->nodes[i];
    int n = sam->madof[node] - sam->madof[node - 1];
    if (n > nNDof) nNDof = n;
    X[i] = xyz[3 * (node - 1)];
    Y[i] = xyz[3 * (node - 1) + 1];
    Z[i] = xyz[3 * (node - 1) + 2];
  }
  if (nNDof > 6) return -1;
  /* calcElmCoordSystem with the position matrix given explicitly (:550-555) */
  ierr = orc_shell_element_axes(nElNodes, X, Y, Z, V1, V2, V3);
  if (ierr != 0) return ierr;
  for (int k = 0; k < 3; k++) { T_el[0 + 3 * k] = V1[k]; T_el[1 + 3 * k] = V2[k]; T_el[2 + 3 * k] = V3[k]; }
#define TEL(i, j) T_el[((i)-1) + 3 * ((j)-1)]
#define POS(i, j) ros->rpos[((i)-1) + 3 * ((j)-1)]
  /* c = matmul(T_el(1:2,:),posInGl(:,1:2)), 2x2 column-major */
  for (int i = 1; i <= 2; i++)
    for (int j = 1; j <= 2; j++)
      c[(i - 1) + 2 * (j - 1)] = TEL(i, 1) * POS(1, j) + TEL(i, 2) * POS(2, j) + TEL(i, 3) * POS(3, j);
#define CC(i, j) c[((i)-1) + 2 * ((j)-1)]
#define TE(i, j) Teps[((i)-1) + 3 * ((j)-1)]
  TE(1, 1) = CC(1, 1) * CC(1, 1);
  TE(1, 2) = CC(2, 1) * CC(2, 1);
  TE(1, 3) = CC(1, 1) * CC(2, 1);
  TE(2, 1) = CC(1, 2) * CC(1, 2);
  TE(2, 2) = CC(2, 2) * CC(2, 2);
  TE(2, 3) = CC(1, 2) * CC(2, 2);
  TE(3, 1) = 2.0 * CC(1, 1) * CC(1, 2);
  TE(3, 2) = 2.0 * CC(2, 2) * CC(2, 1);
  TE(3, 3) = CC(1, 1) * CC(2, 2) + CC(1, 2) * CC(2, 1);
  if (nElNodes == 3)
    orc_strain_disp_cst(nNDof, X, Y, Z, T_el, ros->zpos, B_el);
  else
    orc_strain_disp_quad4(nNDof, X, Y, Z, T_el, 0.0, 0.0, ros->zpos, B_el);
  /* transform to global DOF directions, compressing mixed 3/6-DOF nodes (:697-707) */
  {
    int idof = 1, jdof = 1;
    for (int i = 0; i < nElNodes; i++) {
      int node = ros->nodes[i];
      int n = sam->madof[node] - sam->madof[node - 1];
      for (int j = 0; j <= n - 1; j += 3) {
        double tmp[9];
        for (int a = 0; a < 3; a++)
          for (int b = 1; b <= 3; b++) {
            double s = 0.0;
            for (int k = 1; k <= 3; k++) s += B_el[a + 3 * (jdof + j + k - 2)] * TEL(k, b);
            tmp[a + 3 * (b - 1)] = s;
          }
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++) B_el[a + 3 * (idof + j + b - 1)] = tmp[a + 3 * b];
      }
      for (int d = 0; d < n; d++) rows[idof - 1 + d] = sam->madof[node - 1] + d;
      idof += n;
      jdof += nNDof;
    }
    nElDof = idof - 1;
  }
  for (int i = 0; i < nElDof; i++)
    for (int j = 1; j <= 3; j++)
      bscr[(j - 1) + 3 * i] = TE(j, 1) * B_el[3 * i] + TE(j, 2) * B_el[1 + 3 * i] + TE(j, 3) * B_el[2 + 3 * i];
  *nElDof_out = nElDof;
  return 0;
#undef TEL
#undef POS
#undef CC
#undef TE
}

/* ElDispFromSupElDisp (displacementModule.f90:1096-1202) for ONE rosette followed by
 * Bcart = matmul(bscr,H_el) (strainRosetteModule.f90:727).  Bcart: 3 x ndim column-major.
 * NB (kept as in the reference): the unit response of external DOF i is put at equation meqn2(i),
 * while calcIntDisplacements associates finit(i) with meqn2(dofPosIn2(i)); the two agree whenever
 * meqn2 lists the external DOFs in nodal order (always the case for reducer output). */
int orc_rosette_bcart(const orc_rosette *ros, const orc_sam *sam, const double *xyz,
                      const double *Bmat, const double *Emat, double *Bcart)
{
  const int ndim = sam->ndof2 + sam->ngen, neq = sam->neq;
  double bscr[3 * 24];
  int rows[24], nElDof = 0, ierr;
  double *work, *sv;
  ierr = orc_rosette_bscr(ros, sam, xyz, bscr, rows, &nElDof);
  if (ierr != 0) return ierr;
  work = (double *)calloc((size_t)neq + 1, sizeof(double));
  sv = (double *)calloc((size_t)sam->ndof, sizeof(double));
  for (int c = 1; c <= ndim; c++) {
    const double *v1 = NULL;
    if (c <= sam->ndof2) {
      if (sam->ndof1 > 0) v1 = Bmat + (size_t)(sam->dofPosIn2[c - 1] - 1) * sam->ndof1;
    } else
      v1 = Emat + (size_t)(c - sam->ndof2 - 1) * sam->ndof1;
    if (v1)
      for (int k = 0; k < sam->ndof1; k++) work[sam->meqn1[k] - 1] = v1[k];
    if (c <= sam->ndof2) work[sam->meqn2[c - 1] - 1] = 1.0;
    orc_dis_expand(sam, work, sv);
    if (c <= sam->ndof2) work[sam->meqn2[c - 1] - 1] = 0.0;
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int i = 0; i < nElDof; i++) s += bscr[j + 3 * i] * sv[rows[i] - 1];
      Bcart[j + 3 * (c - 1)] = s;
    }
  }
  free(work);
  free(sv);
  return 0;
}

/* calcRosetteStrains + evaluateStrainGages (strainRosetteModule.f90:225-324).
 * out[ORC_GAGE_NVAL]: [0:3) epsC, [3:6) epsP (max, min, signed abs max), 6 gammaMax, 7 epsVM,
 * 8 alpha1, 9 alphaGamma, [10:13) sigmaC, [13:16) sigmaP, 16 tauMax, 17 sigmaVM,
 * [18:21) epsGage, [21:24) sigGage (unused legs 0). */
void orc_calc_rosette_strains(const double *Bcart, int ndim, const double *finit,
                              const double epsCInit[3], double emod, double nu,
                              const double sigmaC0[3], const double *Tg, int ngage, double *out)
{
  double epsC[3], sigC[3], Cmat[9];
  memset(out, 0, sizeof(double) * ORC_GAGE_NVAL);
  orc_iso_mat2d(emod, nu, Cmat);
  for (int j = 0; j < 3; j++) {
    double s = 0.0;
    for (int c = 0; c < ndim; c++) s += Bcart[j + 3 * c] * finit[c];
    epsC[j] = s + epsCInit[j];
  }
  for (int j = 0; j < 3; j++)
    sigC[j] = Cmat[j] * epsC[0] + Cmat[j + 3] * epsC[1] + Cmat[j + 6] * epsC[2] + sigmaC0[j];
  for (int i = 0; i < ngage && i < 3; i++) {
    out[18 + i] = Tg[3 * i] * epsC[0] + Tg[3 * i + 1] * epsC[1] + Tg[3 * i + 2] * epsC[2];
    out[21 + i] = Tg[3 * i] * sigC[0] + Tg[3 * i + 1] * sigC[1] + Tg[3 * i + 2] * sigC[2];
  }
  for (int j = 0; j < 3; j++) { out[j] = epsC[j]; out[10 + j] = sigC[j]; }
  orc_principle_strains2d(epsC, &out[3], &out[4], &out[6], &out[8], &out[9]);
  out[5] = fabs(out[3]) > fabs(out[4]) ? out[3] : out[4];
  out[7] = sqrt(out[3] * out[3] + out[4] * out[4] - out[3] * out[4]);
  {
    double a1, a2;
    orc_principle_stresses2d(sigC, &out[13], &out[14], &out[16], &a1, &a2);
  }
  out[15] = fabs(out[13]) > fabs(out[14]) ? out[13] : out[14];
  out[17] = sqrt(out[13] * out[13] + out[14] * out[14] - out[13] * out[14]);
}
