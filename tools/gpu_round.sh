#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, ncu launch list + one full capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh TAG
TAG=${1:-r01x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k1_expand|k2_shell_vm' -s 6 -c 4 \
    -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json
