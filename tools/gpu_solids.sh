#!/bin/bash
# GPU visit for the solid kernels: parity tests of the step-lane kernels, A/B lines for HEX20 and config 3, ncu captures.
TAG=${1:-r4a}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "hex20 or tet10 or solid or wedg" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 600 python tools/bench_configs.py hex20 > $O/${TAG}_hex20.json 2> $O/${TAG}_hex20.err
timeout 900 python tools/bench_configs.py c3 --curved surface > $O/${TAG}_c3.json 2> $O/${TAG}_c3.err
timeout 900 python tools/bench_configs.py c3 --curved all > $O/${TAG}_c3_all.json 2>> $O/${TAG}_c3.err
cut -c1-900 $O/${TAG}_hex20.json $O/${TAG}_c3.json $O/${TAG}_c3_all.json
if [ "$2" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k2_hex20_steplane' -s 3 -c 1 \
    -o $O/${TAG}_hex20 -f python tools/bench_configs.py hex20 --steps 2 > $O/${TAG}_ncu_hex20.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_tet10_steplane' -s 6 -c 2 \
    -o $O/${TAG}_tet10 -f python tools/bench_configs.py c3 --curved surface --steps 2 > $O/${TAG}_ncu_tet10.log 2>&1
fi
