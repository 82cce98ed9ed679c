// bin/fedem_fpp: main() of the strain coat recovery program (src/vpmStress/fppmain.C:15-72) over the library's
// initSolverArgs + solveFpp; the option table lives in the library (fsr_fpp_define_options).
#include "../../include/fedem_b200.h"

int main(int argc, char** argv)
{
  initSolverArgs(argc, argv);
  return solveFpp();
}
