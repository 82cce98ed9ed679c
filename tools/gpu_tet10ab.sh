#!/bin/bash
# A/B of step-lane TET10 kernel variants on config 3 (R4 visit; the cp.async-staged and 96-register variants it switched on were
# measured slower and removed again -- see DESIGN.md section 10; kept as the record of how the numbers were taken).
TAG=${1:-R4ab}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_lane" > $O/${TAG}_pytest.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest.log
for cur in surface all; do
  timeout 300 python tools/bench_configs.py c3 --curved $cur > $O/${TAG}_c3_${cur}_default.json 2>&1
  FSR_TET10_STAGED=1 timeout 300 python tools/bench_configs.py c3 --curved $cur > $O/${TAG}_c3_${cur}_staged.json 2>&1
done
FSR_TET10_MINB=5 timeout 300 python tools/bench_configs.py c3 --curved surface > $O/${TAG}_c3_surface_minb5.json 2>&1
tail -2 $O/${TAG}_pytest.log
for f in $O/${TAG}_c3_*.json; do echo $f; python - "$f" <<'P'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: continue
    print("  value %.4e  ms/step %.3f  k2 %.3f  k1 %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d.get("k1_ms", 0)))
P
done
