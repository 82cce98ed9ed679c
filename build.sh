#!/bin/bash
# Builds libfedem_b200.so (sm_100a only) in-tree and the CPU checker under oracle/.
set -e
cd "$(dirname "$0")"
SRC=fedem_solvers_b200/csrc
OUT=fedem_solvers_b200/lib
mkdir -p $OUT build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude"
pids=()
for f in api k1_expand k2_shell k2_solid k2_linsolid k2_hex20 k2_thickshell k2_beam k2_full k3_fatigue k3_gage io_files io_frs io_ftl io_rdb io_fsi stress_driver solver_state sharded; do
  if [ ! -f build/$f.o ] || [ $SRC/$f.cu -nt build/$f.o ] || [ $SRC/common.cuh -nt build/$f.o ] || [ $SRC/io_tagged.cuh -nt build/$f.o ] || [ $SRC/fatigue_core.cuh -nt build/$f.o ] || [ $SRC/invariants.cuh -nt build/$f.o ] || [ $SRC/cmdline.hpp -nt build/$f.o ] || [ include/fedem_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS -Xptxas -v -c $SRC/$f.cu -o build/$f.o 2> build/$f.ptxas.log &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p || { cat build/*.ptxas.log | grep -E "error" ; exit 1; }; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libfedem_b200.so build/api.o build/k1_expand.o build/k2_shell.o build/k2_solid.o build/k2_linsolid.o build/k2_hex20.o build/k2_thickshell.o build/k2_beam.o build/k2_full.o build/k3_fatigue.o build/k3_gage.o build/io_files.o build/io_frs.o build/io_ftl.o build/io_rdb.o build/io_fsi.o build/stress_driver.o build/solver_state.o build/sharded.o -lcudart -ldl
mkdir -p fedem_solvers_b200/bin
g++ -O2 -o fedem_solvers_b200/bin/fedem_stress $SRC/stress_main.cpp -L$OUT -lfedem_b200 -Wl,-rpath,'$ORIGIN/../lib'
g++ -O2 -o fedem_solvers_b200/bin/fedem_gage $SRC/gage_main.cpp -L$OUT -lfedem_b200 -Wl,-rpath,'$ORIGIN/../lib'
g++ -O2 -o fedem_solvers_b200/bin/fedem_modes $SRC/modes_main.cpp -L$OUT -lfedem_b200 -Wl,-rpath,'$ORIGIN/../lib'
g++ -O2 -o fedem_solvers_b200/bin/fedem_fpp $SRC/fpp_main.cpp -L$OUT -lfedem_b200 -Wl,-rpath,'$ORIGIN/../lib'
make -s -C oracle all
echo "built $OUT/libfedem_b200.so"
