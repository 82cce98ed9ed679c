// modes_main.cpp -- the fedem_modes executable: same shape as the reference's main() (src/vpmStress/modesmain.C:15-56):
// initialise the command-line parser, define the options, run.  Everything lives in libfedem_b200.so
// (csrc/stress_driver.cu) under the reference's exported names.
extern "C" {
void initSolverArgs(int argc, char** argv);
int solveModes(void);
}

int main(int argc, char** argv)
{
  initSolverArgs(argc, argv);
  return solveModes();
}
