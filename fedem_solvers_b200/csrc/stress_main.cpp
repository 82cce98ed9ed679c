// stress_main.cpp -- the fedem_stress executable: same shape as the reference's main()
// (src/vpmStress/stressmain.C:16-83): initialise the command-line parser, define the options, run.
// Everything lives in libfedem_b200.so (csrc/stress_driver.cu) under the reference's exported names.
extern "C" {
void initSolverArgs(int argc, char** argv);
int solveStress(void);
}

int main(int argc, char** argv)
{
  initSolverArgs(argc, argv);
  return solveStress();
}
