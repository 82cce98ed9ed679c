/* c_abi_driver.c -- TEST: a plain C caller of libfedem_b200.so through include/fedem_b200.h only
 * (the same symbols and struct layouts the Fortran interface module binds).  With a GPU it runs a
 * 2x2-quad plate through create -> set_recovery -> recover -> envelope and prints the numbers; with
 * no GPU it must fail loudly with FSR_ERR_CUDA.  Exit code 0 in both cases when behaviour is right. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fedem_b200.h"

int main(void)
{
  enum { NX = 2, NN = 9, NEL = 4, NDOF = 54, NEXT = 6, NINT = 48, NGEN = 2, NDIM = 8, NSTEP = 5 };
  int madof[NN + 1], msc[NDOF], mpmnpc[NEL + 1], mmnpc[4 * NEL], melcon[NEL], meqn[NDOF], meqn1[NINT], meqn2[NEXT];
  int mpmceq[1] = {1}, mmceq[1] = {0};
  double ttcc[1] = {0}, xyz[3 * NN], emod[NEL], rny[NEL], thk[NEL];
  double *B = calloc((size_t)NINT * NEXT, sizeof(double)), *E = calloc((size_t)NINT * NGEN, sizeof(double));
  double Q[NDIM * NSTEP];
  int i, j, e = 0, n1 = 0, n2 = 0;
  fsr_sam sam;
  fsr_elmdata elm;
  fsr_options opt;
  fsr_part *part = NULL;
  int rc;

  for (i = 0; i <= NN; i++) madof[i] = 1 + 6 * i;
  for (j = 0; j < 3; j++)
    for (i = 0; i < 3; i++) {
      xyz[3 * (3 * j + i)] = 0.5 * i + (i == 1 && j == 1 ? 0.03 : 0.0);
      xyz[3 * (3 * j + i) + 1] = 0.5 * j;
      xyz[3 * (3 * j + i) + 2] = 0.0;
    }
  for (j = 0; j < NX; j++)
    for (i = 0; i < NX; i++, e++) {
      int n = 3 * j + i + 1;
      mpmnpc[e] = 1 + 4 * e;
      mmnpc[4 * e] = n; mmnpc[4 * e + 1] = n + 1; mmnpc[4 * e + 2] = n + 4; mmnpc[4 * e + 3] = n + 3;
      melcon[e] = 24; emod[e] = 2.1e11; rny[e] = 0.3; thk[e] = 0.01;
    }
  mpmnpc[NEL] = 1 + 4 * NEL;
  for (i = 0; i < NDOF; i++) {       /* node 1 is the external node */
    msc[i] = i < 6 ? 2 : 1;
    meqn[i] = i + 1;
    if (i < 6) meqn2[n2++] = i + 1; else meqn1[n1++] = i + 1;
  }
  for (i = 0; i < NINT; i++) {
    for (j = 0; j < NEXT; j++) B[i + (size_t)NINT * j] = 0.01 * cos(0.3 * i + j);
    for (j = 0; j < NGEN; j++) E[i + (size_t)NINT * j] = 0.02 * sin(0.2 * i + 2 * j);
  }
  for (i = 0; i < NDIM * NSTEP; i++) Q[i] = 1e-3 * sin(0.7 * i);

  memset(&sam, 0, sizeof(sam));
  sam.nnod = NN; sam.nel = NEL; sam.ndof = NDOF; sam.ndof1 = NINT; sam.ndof2 = NEXT; sam.ngen = NGEN;
  sam.neq = NDOF; sam.nceq = 0; sam.nmmnpc = 4 * NEL; sam.nmmceq = 0;
  sam.madof = madof; sam.msc = msc; sam.mpmnpc = mpmnpc; sam.mmnpc = mmnpc; sam.melcon = melcon;
  sam.mpmceq = mpmceq; sam.mmceq = mmceq; sam.ttcc = ttcc; sam.meqn = meqn; sam.meqn1 = meqn1; sam.meqn2 = meqn2;
  memset(&elm, 0, sizeof(elm));
  elm.xyz = xyz; elm.emod = emod; elm.rny = rny; elm.thk = thk;
  memset(&opt, 0, sizeof(opt));

  rc = fsr_part_create(&part, &sam, &elm, &opt);
  if (rc == FSR_ERR_CUDA) {
    printf("no usable B200: %s\n", fsr_last_error());
    return strlen(fsr_last_error()) > 0 ? 0 : 1; /* failing loudly is the required behaviour */
  }
  if (rc < 0) { printf("fsr_part_create failed: %s\n", fsr_last_error()); return 1; }
  if (fsr_set_recovery(part, B, NINT, E, NINT) < 0) { printf("%s\n", fsr_last_error()); return 1; }
  {
    int npts = fsr_num_result_points(part);
    double *vm = malloc(sizeof(double) * (size_t)npts * NSTEP), *mx = malloc(sizeof(double) * npts), *mn = malloc(sizeof(double) * npts);
    if (npts != 8 * NEL || fsr_ndim(part) != NDIM) { printf("bad sizes %d %d\n", npts, fsr_ndim(part)); return 1; }
    if (fsr_recover(part, Q, NDIM, NSTEP, vm) < 0) { printf("%s\n", fsr_last_error()); return 1; }
    if (fsr_get_envelope(part, mx, mn) < 0) { printf("%s\n", fsr_last_error()); return 1; }
    for (i = 0; i < npts; i++) {
      double hi = 0.0, lo = 1e308;
      for (j = 0; j < NSTEP; j++) { double v = vm[(size_t)j * npts + i]; if (v > hi) hi = v; if (v < lo) lo = v; }
      if (hi != mx[i] || lo != mn[i] || !(hi > 0.0)) { printf("envelope mismatch at point %d\n", i); return 1; }
    }
    printf("C ABI driver OK: %d result points, vm[0][0] = %.6e, launches = %lld\n", npts, vm[0], fsr_kernel_launches(0));
    free(vm); free(mx); free(mn);
  }
  fsr_part_destroy(part);
  free(B); free(E);
  return 0;
}
