#!/bin/bash
# GPU visit for the in-plane form of flat shell regions: parity, bench with / without it, ncu of K1 + the quad kernel.
TAG=${1:-r2m}
O=gpurun_out
mkdir -p $O
NOSEC="--no-secondary --no-cpu-baseline"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 600 python bench.py $NOSEC > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
FSR_QUAD_PLANAR=0 timeout 600 python bench.py $NOSEC --no-parity > $O/${TAG}_bench_noplanar.json 2> $O/${TAG}_bench_noplanar.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k1_expand|k2_quad_planar' -s 6 -c 4 \
    -o $O/${TAG}_k1_k2 -f python bench.py --steps 2 --warmup 3 $NOSEC --no-parity > $O/${TAG}_ncu_k1_k2.log 2>&1
tail -3 $O/${TAG}_pytest.log; cut -c1-900 $O/${TAG}_bench.json; echo; cut -c1-300 $O/${TAG}_bench_noplanar.json
