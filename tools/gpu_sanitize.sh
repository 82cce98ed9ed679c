#!/bin/bash
# compute-sanitizer over the small GPU cases: memcheck of smoke() and of the fatigue / gage / parity tests (out-of-bounds and
# misaligned accesses in every kernel they launch), racecheck of the kernels that exchange data through shared memory.
# Usage (from the repo root, under gpurun):  bash tools/gpu_sanitize.sh TAG
TAG=${1:-R4}
O=gpurun_out
mkdir -p $O
CS="compute-sanitizer --error-exitcode 9 --print-limit 20"
timeout 150 $CS --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_memcheck_smoke.log 2>&1; echo "rc=$?" >> $O/${TAG}_memcheck_smoke.log
timeout 240 $CS --tool memcheck python -m pytest tests/test_gpu_fatigue.py tests/test_gpu_gage.py -m gpu -x -q > $O/${TAG}_memcheck_k3.log 2>&1; echo "rc=$?" >> $O/${TAG}_memcheck_k3.log
timeout 300 $CS --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1 or tet10 or hex20 or inplane or triangles or large_reduced or thick or edge_sizes or wedg15" > $O/${TAG}_memcheck_parity.log 2>&1; echo "rc=$?" >> $O/${TAG}_memcheck_parity.log
timeout 150 $CS --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_racecheck_smoke.log 2>&1; echo "rc=$?" >> $O/${TAG}_racecheck_smoke.log
timeout 200 $CS --tool racecheck python -m pytest tests/test_gpu_fatigue.py -m gpu -x -q > $O/${TAG}_racecheck_k3.log 2>&1; echo "rc=$?" >> $O/${TAG}_racecheck_k3.log
for f in $O/${TAG}_memcheck_*.log $O/${TAG}_racecheck_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $f | tail -4; done
