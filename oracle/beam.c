/* beam.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): 2-node beam section forces STR11.
 * Placeholder until the BEAM31/BELS31 restatement lands (src/vpmStress/elStressModule.f90:
 * 402-515, src/Femlib/beam.f:11-166,619-803, src/Femlib/beamaux.f:48-260). */
#include "oracle.h"
int orc_str11(const double *beam, const double ev[12], double SF[12])
{
  (void)beam; (void)ev;
  for (int i = 0; i < 12; i++) SF[i] = 0.0;
  return 1;
}
