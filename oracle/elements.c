/* elements.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): element dispatch and the
 * per-step element loop.  Follows src/vpmStress/elStressModule.f90:58-99 (extractEV),
 * :129-253 (ElStress incl. the tensorial-shear conversion), src/vpmStress/stressRoutines.f90:
 * 169-342 (calcStresses element loop, hugeVal for failed elements, in-core vms order),
 * src/vpmStress/stress.f90:357-435 (time loop), strainCoatModule.f90:159-166,410-420
 * (envelope semantics: running max initialised to 0, running min to hugeVal). */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* elStressModule.f90:58-99.  Returns NEDOF (negative on overflow). */
int orc_extract_ev(int iel, const orc_sam *sam, const double *sv, double *ev, int evsize)
{
  int nedof = 0, nndof;
  int t = sam->melcon[iel - 1];
  if (t == 11 || (t >= 21 && t <= 24) || t == 31 || t == 32)
    nndof = 6;
  else if (t >= 41 && t <= 46)
    nndof = 3;
  else
    return 0;
  for (int ip = sam->mpmnpc[iel - 1]; ip <= sam->mpmnpc[iel] - 1; ip++) {
    int in = sam->mmnpc[ip - 1];
    int js = sam->madof[in - 1];
    int nd = sam->madof[in] - js;
    if (nd > nndof) nd = nndof;
    int iedof = nedof + 1;
    nedof = nedof + nd;
    /* The reference tests NEDOF < size(EV) with EV(60) (stressRoutines.f90:86,93): for the one
     * element type with exactly 60 DOFs, HEX20, the 20th node is then NOT copied and keeps whatever
     * the previous element left in EV(58:60).  That stale read cannot be a parity target (it depends
     * on the element visited before); DELIBERATE DEVIATION: '<=' here and in the CUDA path, i.e. all
     * 20 nodes are used.  Identical for every other element type (NEDOF < 60).  See DESIGN.md 9. */
    if (nedof <= evsize)
      for (int k = 0; k < nd; k++) ev[iedof - 1 + k] = sv[js - 1 + k];
  }
  if (nedof > evsize) nedof = evsize - nedof;
  return nedof;
}

static void get_coor(int iel, const orc_sam *sam, const orc_elmdata *ed, int n, double *x,
                     double *y, double *z)
{
  int ip0 = sam->mpmnpc[iel - 1];
  for (int k = 0; k < n; k++) {
    int in = sam->mmnpc[ip0 - 1 + k];
    x[k] = ed->xyz[3 * (in - 1)];
    y[k] = ed->xyz[3 * (in - 1) + 1];
    z[k] = ed->xyz[3 * (in - 1) + 2];
  }
}

/* elStressModule.f90:129-253.  S = stress resultants / section forces (6 x nenod),
 * Sigma/Epsil = (ncmp x nstrp).  Returns ierr (0 ok, >0 element failed, <0 fatal). */
int orc_el_stress(int iel, int ieltyp, const orc_sam *sam, const orc_elmdata *ed,
                  double *V, double *S, double *Sigma, double *Epsil, int *nenod, int *nstrp)
{
  double x[20], y[20], z[20], thk[8], SS[24];
  int ierr = 0;
  *nenod = 0;
  *nstrp = 0;
  if (ed->elmid && ed->elmid[iel - 1] < 1) return 0;

  switch (ieltyp) {
  case 11:
    *nenod = 2;
    *nstrp = 0;
    if (!ed->beam) return 1;
    ierr = orc_str11(ed->beam + (size_t)ORC_NBEAM * (iel - 1), V, S);
    Sigma[0] = 0.0;
    Epsil[0] = 0.0;
    break;
  case 21: /* FFT3 with the default -fftStressForm 1: STR21 runs the statements of STR23 (elStressModule.f90:559-562,586-587) */
  case 23:
    *nenod = 3;
    *nstrp = 6;
    get_coor(iel, sam, ed, 3, x, y, z);
    thk[0] = thk[1] = thk[2] = ed->thk[iel - 1];
    if (ieltyp == 21 && orc_get_fft_stress_form() != 1) {
      ierr = orc_str21_legacy(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], thk, V, S, SS, Sigma, Epsil);
      break;
    }
    ierr = orc_str23(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], thk, V, S, SS, Sigma,
                     Epsil);
    break;
  case 22: /* FFQ4: STR22 = pMatStiff + STR22a(nGauss) for -ffqStressForm 1 / 2 (:675-686,717-722); 2 = the statements of STR24 */
    *nenod = 4;
    *nstrp = 8;
    get_coor(iel, sam, ed, 4, x, y, z);
    thk[0] = thk[1] = thk[2] = thk[3] = ed->thk[iel - 1];
    ierr = orc_str22(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], thk, V, S, SS, Sigma, Epsil);
    break;
  case 24:
    *nenod = 4;
    *nstrp = 8;
    get_coor(iel, sam, ed, 4, x, y, z);
    thk[0] = thk[1] = thk[2] = thk[3] = ed->thk[iel - 1];
    ierr = orc_str24(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], thk, V, S, SS, Sigma,
                     Epsil);
    break;
  case 31:
    *nenod = 6;
    *nstrp = 12;
    get_coor(iel, sam, ed, 6, x, y, z);
    for (int k = 0; k < 6; k++) thk[k] = ed->thk[iel - 1];
    memset(S, 0, sizeof(double) * 36);   /* "Nodal stress resultants ... maybe later", elStressModule.f90:1174-1176 */
    ierr = orc_str31(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], thk, V, Sigma, Epsil);
    break;
  case 32:
    *nenod = 8;
    *nstrp = 16;
    get_coor(iel, sam, ed, 8, x, y, z);
    for (int k = 0; k < 8; k++) thk[k] = ed->thk[iel - 1];
    memset(S, 0, sizeof(double) * 48);
    ierr = orc_str32(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], thk, V, Sigma, Epsil);
    break;
  case 41:
    *nenod = 10;
    *nstrp = 10;
    get_coor(iel, sam, ed, 10, x, y, z);
    ierr = orc_str41(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], 0, V, Sigma, Epsil);
    break;
  case 43:
    *nenod = 20;
    *nstrp = 20;
    get_coor(iel, sam, ed, 20, x, y, z);
    ierr = orc_str43(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], 0, V, Sigma, Epsil);
    break;
  case 42:
    *nenod = 15;
    *nstrp = 15;
    get_coor(iel, sam, ed, 15, x, y, z);
    ierr = orc_str42(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], 0, V, Sigma, Epsil);
    break;
  case 44:
    *nenod = 8;
    *nstrp = 8;
    get_coor(iel, sam, ed, 8, x, y, z);
    ierr = orc_str44(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], 0, V, Sigma, Epsil);
    break;
  case 45:
    *nenod = 4;
    *nstrp = 4;
    get_coor(iel, sam, ed, 4, x, y, z);
    ierr = orc_str45(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], V, Sigma, Epsil);
    break;
  case 46:
    *nenod = 6;
    *nstrp = 6;
    get_coor(iel, sam, ed, 6, x, y, z);
    ierr = orc_str46(x, y, z, ed->emod[iel - 1], ed->rny[iel - 1], 0, V, Sigma, Epsil);
    break;
  default:
    return 0; /* silently ignore all other element types */
  }
  if (ierr != 0) return ierr;

  for (int n = 1; n <= *nstrp; n++) {
    if (ieltyp >= 21 && ieltyp <= 24)
      Epsil[3 * n - 1] = Epsil[3 * n - 1] * 0.5;
    else if (ieltyp == 31 || ieltyp == 32 || (ieltyp >= 41 && ieltyp <= 46)) {
      Epsil[6 * n - 3] = Epsil[6 * n - 3] * 0.5;
      Epsil[6 * n - 2] = Epsil[6 * n - 2] * 0.5;
      Epsil[6 * n - 1] = Epsil[6 * n - 1] * 0.5;
    }
  }
  return 0;
}

static int ncomp1(int t) /* stressRoutines.f90:393-400 */
{
  if (t == 11) return 1;
  if (t >= 21 && t <= 24) return 3;
  if (t == 31 || t == 32) return 6;
  if (t >= 41 && t <= 46) return 6;
  return 0;
}
static int nstrp_of(int t) /* elStressModule.f90:159-229 */
{
  switch (t) {
  case 11: return 0;
  case 21: case 23: return 6;
  case 22: case 24: return 8;
  case 31: return 12;
  case 32: return 16;
  case 41: return 10;
  case 42: return 15;
  case 43: return 20;
  case 44: return 8;
  case 45: return 4;
  case 46: return 6;
  }
  return 0;
}

/* Running offsets of the result points of each element in processing (SAM) order; elements
 * that are skipped (elmid < 1) or have no stress points contribute nothing
 * (stressRoutines.f90:257,324-331).  off has nel+1 entries; returns total number of points. */
int orc_result_point_offsets(const orc_sam *sam, const int *elmid, int *off)
{
  int n = 0;
  for (int iel = 1; iel <= sam->nel; iel++) {
    off[iel - 1] = n;
    if (elmid && elmid[iel - 1] < 1) continue;
    n += nstrp_of(sam->melcon[iel - 1]);
  }
  off[sam->nel] = n;
  return n;
}

/* One time step of stressRoutines.f90:169-342 with all eight derived measures requested.
 * resmat : [8 x npts] column-major per point: vmStress, maxP, minP, maxShear (stress), then
 *          the same four for strain (stressRoutines.f90:273-287); hugeVal for failed elements.
 * stress/strain : [6 x npts] (ncmp leading entries used), may be NULL.
 * sres   : [24 x nel] stress resultants SR(6, node 1..4) of the thin shells / beam section forces
 *          SF(6,2), may be NULL.
 * Returns number of failed elements. */
int orc_calc_stresses(const orc_sam *sam, const orc_elmdata *ed, const double *sv,
                      const int *ptoff, double *resmat, double *stress, double *strain,
                      double *sres, int nthreads)
{
  int nfail = 0;
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : nfail)
#endif
  for (int iel = 1; iel <= sam->nel; iel++) {
    double EV[60], SR[60], Stress[300], Strain[300], rm[8][50];
    int nenod = 0, nstrp = 0, lerr;
    int t = sam->melcon[iel - 1];
    int ncmp = ncomp1(t);
    memset(SR, 0, sizeof(SR));
    memset(Stress, 0, sizeof(Stress));
    memset(Strain, 0, sizeof(Strain));
    lerr = orc_extract_ev(iel, sam, sv, EV, 60);
    if (lerr > 0)
      lerr = orc_el_stress(iel, t, sam, ed, EV, SR, Stress, Strain, &nenod, &nstrp);
    if (sres && nenod > 0)
      for (int k = 0; k < 24; k++) sres[(size_t)24 * (iel - 1) + k] = lerr > 0 ? ORC_HUGE : SR[k];
    if (ncmp > 0 && nstrp > 0) {
      int p0 = ptoff[iel - 1];
      if (lerr > 0) {
        nfail++;
        for (int p = 0; p < nstrp; p++) {
          for (int k = 0; k < 8; k++) resmat[(size_t)8 * (p0 + p) + k] = ORC_HUGE;
          if (stress) for (int k = 0; k < 6; k++) stress[(size_t)6 * (p0 + p) + k] = ORC_HUGE;
          if (strain) for (int k = 0; k < 6; k++) strain[(size_t)6 * (p0 + p) + k] = ORC_HUGE;
        }
      } else {
        orc_calc_von_mises(Stress, ncmp, nstrp, rm[0]);
        orc_calc_principal_vals(Stress, ncmp, nstrp, rm[1], rm[2], rm[3]);
        orc_calc_von_mises(Strain, ncmp, nstrp, rm[4]);
        orc_calc_principal_vals(Strain, ncmp, nstrp, rm[5], rm[6], rm[7]);
        for (int p = 0; p < nstrp; p++) {
          for (int k = 0; k < 8; k++) resmat[(size_t)8 * (p0 + p) + k] = rm[k][p];
          if (stress) {
            for (int k = 0; k < 6; k++) stress[(size_t)6 * (p0 + p) + k] = 0.0;
            for (int k = 0; k < ncmp; k++) stress[(size_t)6 * (p0 + p) + k] = Stress[ncmp * p + k];
          }
          if (strain) {
            for (int k = 0; k < 6; k++) strain[(size_t)6 * (p0 + p) + k] = 0.0;
            for (int k = 0; k < ncmp; k++) strain[(size_t)6 * (p0 + p) + k] = Strain[ncmp * p + k];
          }
        }
      }
    }
  }
  return nfail;
}

/* The reference's time loop (stress.f90:361-435) over a reduced history Q (ndim x nsteps,
 * column-major: column = [finit(1:ndof2); vg(1:ngen)] of one step), keeping the reference's
 * loop structure: per step two column-AXPY mat-vecs, then every element rebuilt from its
 * coordinates.  vm_hist [nsteps x npts] step-major (may be NULL); env_max/env_min [npts]
 * running envelopes of von Mises (strainCoatModule.f90:159-166,410-420), may be NULL.
 * This is also what bench.py times as the CPU baseline. */
int orc_recover_history(const orc_sam *sam, const orc_elmdata *ed, const double *Bmat,
                        const double *Emat, const double *Q, int nsteps, const int *ptoff,
                        double *vm_hist, double *env_max, double *env_min, int nthreads)
{
  const int ndim = sam->ndof2 + sam->ngen;
  const int npts = ptoff[sam->nel];
  double *work = (double *)malloc(sizeof(double) * ((size_t)sam->neq + sam->ndof1 + sam->ndof2));
  double *sv = (double *)malloc(sizeof(double) * (size_t)sam->ndof);
  double *resmat = (double *)malloc(sizeof(double) * 8 * (size_t)(npts > 0 ? npts : 1));
  if (!work || !sv || !resmat) { free(work); free(sv); free(resmat); return -1; }
  {
    extern int orc_threads;
    orc_threads = nthreads > 0 ? nthreads : 1;
  }
  if (env_max) for (int p = 0; p < npts; p++) env_max[p] = 0.0;
  if (env_min) for (int p = 0; p < npts; p++) env_min[p] = ORC_HUGE;
  for (int s = 0; s < nsteps; s++) {
    const double *q = Q + (size_t)ndim * s;
    orc_calc_int_displacements(sam, Bmat, Emat, q, q + sam->ndof2, work, sv);
    orc_calc_stresses(sam, ed, sv, ptoff, resmat, NULL, NULL, NULL, nthreads);
    for (int p = 0; p < npts; p++) {
      double vm = resmat[(size_t)8 * p];
      if (vm_hist) vm_hist[(size_t)npts * s + p] = vm;
      if (env_max && vm > env_max[p]) env_max[p] = vm;
      if (env_min && vm < env_min[p]) env_min[p] = vm;
    }
  }
  free(work); free(sv); free(resmat);
  return 0;
}
