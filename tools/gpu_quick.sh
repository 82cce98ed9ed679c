#!/bin/bash
# Quick GPU visit: the GPU parity tests only.  Usage (under gpurun): bash tools/gpu_quick.sh TAG [pytest args]
TAG=${1:-q}; shift
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -25 $O/${TAG}_pytest.log
