/* expand.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): reduced history -> nodal displacements.
 * Follows src/vpmStress/displacementModule.f90:931-1024 (calcIntDisplacements), :1226-1259
 * (disExpand), src/vpmUtilities/diskMatrixModule.f90:993-1045 (dmMatTimesVec),
 * SAM/src/dscatr.f (DSCATR), src/vpmCommon/samModule.f90:960-970 (dofPosIn2). */
#include "oracle.h"
#include <string.h>

int orc_threads = 1; /* host threads for the bench variant; 1 = the reference's serial loop */

/* samModule.f90:964-970: for the i2-th DOF with status code 2 (in nodal DOF order),
 * dofPosIn2(i2) = findloc(meqn2, meqn(idof)) (1-based position, 0 if absent). */
void orc_dof_pos_in2(int ndof, int ndof2, const int *msc, const int *meqn,
                     const int *meqn2, int *dofPosIn2)
{
  int i2 = 0;
  for (int idof = 1; idof <= ndof; idof++)
    if (msc[idof - 1] == 2) {
      int pos = 0;
      for (int j = 1; j <= ndof2; j++)
        if (meqn2[j - 1] == meqn[idof - 1]) { pos = j; break; }
      if (i2 < ndof2) dofPosIn2[i2] = pos;
      i2++;
    }
}

/* diskMatrixModule.f90:1024-1041: y = A*x or y = y + A*x, one AXPY per column
 * (A column-major, nrows x ncols, all columns in core). */
void orc_mat_times_vec(int nrows, int ncols, const double *A, const double *x,
                       double *y, int do_initialize)
{
  if (nrows < 1 || ncols < 1) return;
  if (do_initialize) memset(y, 0, sizeof(double) * (size_t)nrows);
  /* Row blocks may run on several host threads (bench "all host cores" variant); every y[r]
   * still accumulates its columns in the reference's order i = 1..ncols. */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(orc_threads > 0 ? orc_threads : 1)
#endif
  for (int r0 = 0; r0 < nrows; r0 += 4096) {
    int r1 = r0 + 4096 < nrows ? r0 + 4096 : nrows;
    for (int i = 0; i < ncols; i++) {
      const double *col = A + (size_t)i * (size_t)nrows;
      const double xi = x[i];
      for (int r = r0; r < r1; r++) y[r] = y[r] + col[r] * xi;
    }
  }
}

/* displacementModule.f90:1239-1257 */
void orc_dis_expand(const orc_sam *sam, const double *sveq, double *svdof)
{
  for (int idof = 1; idof <= sam->ndof; idof++) {
    int ieq = sam->meqn[idof - 1];
    int iceq = -ieq;
    if (ieq > 0 && ieq <= sam->neq)
      svdof[idof - 1] = sveq[ieq - 1];
    else if (iceq > 0 && iceq <= sam->nceq) {
      double s = 0.0;
      for (int ip = sam->mpmceq[iceq - 1] + 1; ip <= sam->mpmceq[iceq] - 1; ip++) {
        int m = sam->mmceq[ip - 1];
        if (m > 0 && m <= sam->ndof) {
          int jeq = sam->meqn[m - 1];
          if (jeq > 0 && jeq <= sam->neq) s = s + sam->ttcc[ip - 1] * sveq[jeq - 1];
        }
      }
      svdof[idof - 1] = s;
    } else
      svdof[idof - 1] = 0.0;
  }
}

/* displacementModule.f90:956-1003.  work must hold neq+ndof1+ndof2 doubles.
 * Bmat: ndof1 x ndof2, Emat: ndof1 x ngen, both column-major (as dmOpen reads them). */
int orc_calc_int_displacements(const orc_sam *sam, const double *Bmat, const double *Emat,
                               const double *finit, const double *vg, double *work,
                               double *sv)
{
  const int neq = sam->neq, ndof1 = sam->ndof1, ndof2 = sam->ndof2, ngen = sam->ngen;
  double *sveq = work;
  double *vi = work + neq;
  double *ve = vi + ndof1;

  /* :974-980 extract the external DOFs: ve(dofPosIn2(i)) = finit(i) */
  for (int i = 0; i < ndof2; i++) ve[i] = 0.0;
  for (int i = 0; i < ndof2; i++) ve[sam->dofPosIn2[i] - 1] = finit[i];

  /* :984-989 internal DOFs */
  orc_mat_times_vec(ndof1, ndof2, Bmat, ve, vi, 1);
  if (ngen > 0) orc_mat_times_vec(ndof1, ngen, Emat, vg, vi, 0);

  /* :998-1000 scatter into equation order */
  for (int i = 0; i < neq; i++) sveq[i] = 0.0;
  for (int i = 0; i < ndof1; i++) sveq[sam->meqn1[i] - 1] = vi[i];
  for (int i = 0; i < ndof2; i++) sveq[sam->meqn2[i] - 1] = ve[i];

  /* :1003 */
  orc_dis_expand(sam, sveq, sv);
  return 0;
}
