/* coat.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): the strain coat recovery summary of one result point.
 * Follows src/vpmStress/strainCoatModule.f90: nullifyResults :142-170, calcStrainCoatData :315-480 (updateMax / updateMin,
 * updateBiAxial, updateAngBin with its allocate-on-first-touch bins), calcAngleData :481-547 (useOldRange = .false.),
 * BiAxMean / BiAxStdDev :672-704.  Input: the per-step rosette values of orc_calc_rosette_strains (ORC_GAGE_NVAL per step). */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>

typedef struct { int alloc, nVal; double sigMax, sigMin, epsMax, epsMin; } angbin;

/* env[8]: epsMax, epsMin, sigMax, sigMin, gammaMax, tauMax, vmeMax, vmsMax; summary[6]: sRange(1:2), popAng, angSpd, biaxial mean,
 * biaxial standard deviation; nval_out [angle_bins-1] (may be NULL): hit count of every bin, -1 = never allocated.
 * Returns nBiAxial. */
int orc_coat_summary(const double *values, int nsteps, int angle_bins, double biaxial_gate, double *env, double *summary,
                     int *nval_out)
{
  const double pi_p = 3.141592653589793238;
  const int nBin = angle_bins - 1;
  angbin *bin = (angbin *)calloc((size_t)nBin, sizeof(angbin));
  double epsMax = 0.0, epsMin = ORC_HUGE, sigMax = 0.0, sigMin = ORC_HUGE, gammaMax = 0.0, tauMax = 0.0, vmeMax = 0.0, vmsMax = 0.0;
  double biAxialSum = 0.0, biAxialSqr = 0.0;
  int nBiAxial = 0;
  for (int t = 0; t < nsteps; t++) {
    const double *v = values + (size_t)ORC_GAGE_NVAL * t;
    const double *epsP = v + 3, *sigmaP = v + 13;
    if (epsP[0] > epsMax) epsMax = epsP[0];
    if (epsP[1] < epsMin) epsMin = epsP[1];
    if (sigmaP[0] > sigMax) sigMax = sigmaP[0];
    if (sigmaP[1] < sigMin) sigMin = sigmaP[1];
    if (v[6] > gammaMax) gammaMax = v[6];
    if (v[16] > tauMax) tauMax = v[16];
    if (v[7] > vmeMax) vmeMax = v[7];
    if (v[17] > vmsMax) vmsMax = v[17];
    { /* updateAngBin */
      const double angle = v[8];
      int iAng = (int)lround((angle / pi_p + 0.5) * nBin);
      int jAng = (int)lround((angle / pi_p + 1.0) * nBin);
      if (iAng < 1) iAng = nBin;
      if (jAng > nBin) jAng = jAng - nBin;
      angbin *b = &bin[iAng - 1];
      if (!b->alloc) {
        b->alloc = 1; b->nVal = 1;
        b->sigMax = sigmaP[0]; b->sigMin = sigmaP[0]; b->epsMax = epsP[0]; b->epsMin = epsP[0];
      } else {
        b->nVal = b->nVal + 1;
        b->sigMax = fmax(b->sigMax, sigmaP[0]); b->sigMin = fmin(b->sigMin, sigmaP[0]);
        b->epsMax = fmax(b->epsMax, epsP[0]); b->epsMin = fmin(b->epsMin, epsP[0]);
      }
      b = &bin[jAng - 1];
      if (!b->alloc) {
        b->alloc = 1; b->nVal = 0;
        b->sigMax = sigmaP[1]; b->sigMin = sigmaP[1]; b->epsMax = epsP[1]; b->epsMin = epsP[1];
      } else {
        b->sigMax = fmax(b->sigMax, sigmaP[1]); b->sigMin = fmin(b->sigMin, sigmaP[1]);
        b->epsMax = fmax(b->epsMax, epsP[1]); b->epsMin = fmin(b->epsMin, epsP[1]);
      }
    }
    if (sigmaP[2] > biaxial_gate) { /* updateBiAxial */
      double biaxial;
      if (fabs(sigmaP[0]) > fabs(sigmaP[1])) biaxial = sigmaP[1] / sigmaP[0];
      else biaxial = sigmaP[0] / sigmaP[1];
      biAxialSum = biAxialSum + biaxial;
      biAxialSqr = biAxialSqr + biaxial * biaxial;
      nBiAxial = nBiAxial + 1;
    }
  }
  env[0] = epsMax; env[1] = epsMin; env[2] = sigMax; env[3] = sigMin; env[4] = gammaMax; env[5] = tauMax; env[6] = vmeMax; env[7] = vmsMax;
  if (nval_out) for (int i = 0; i < nBin; i++) nval_out[i] = bin[i].alloc ? bin[i].nVal : -1;
  { /* calcAngleData */
    const double binSize = 180.0 / nBin;
    double sRange[2] = {0.0, 0.0};
    int iGap = 0, mVal = 0, firstGap, maxGap;
    for (int i = 1; i <= nBin; i++) {
      angbin *b = &bin[i - 1];
      if (b->alloc) {
        sRange[0] = fmax(sRange[0], b->sigMax - b->sigMin);
        sRange[1] = fmax(sRange[1], b->epsMax - b->epsMin);
        if (b->nVal > mVal) { iGap = i; mVal = b->nVal; }
        else if (b->nVal == 0) b->alloc = 0;
      }
    }
    summary[0] = sRange[0]; summary[1] = sRange[1];
    summary[2] = iGap * binSize - 90.0;
    firstGap = 0; maxGap = 0; iGap = 0;
    for (int i = 1; i <= nBin; i++) {
      if (bin[i - 1].alloc) {
        if (firstGap == 0) firstGap = i;
        else if (iGap > 0) { if (i - iGap + 1 > maxGap) maxGap = i - iGap + 1; iGap = 0; }
      } else if (iGap == 0 && firstGap > 0)
        iGap = i;
    }
    if (iGap > 0) firstGap = firstGap + nBin - iGap + 1;
    if (firstGap > maxGap) maxGap = firstGap;
    summary[3] = 180.0 - maxGap * binSize;
  }
  summary[4] = biAxialSum / (nBiAxial > 1 ? nBiAxial : 1);
  summary[5] = 0.0;
  {
    double dnum = nBiAxial;
    if (dnum > 1.0) {
      double mean = biAxialSum / dnum, dvar = biAxialSqr / dnum - mean * mean;
      if (dvar > 0.0) summary[5] = sqrt(dvar * dnum / (dnum - 1.0));
    }
  }
  free(bin);
  return nBiAxial;
}
