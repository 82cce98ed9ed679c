#!/bin/bash
# GPU visit R4 (after the step-lane solid kernels and the cheaper K3 damage / pairwise gage post kernel): parity tests, both
# bench arms, launch lists of the bench command and of C5, ncu full captures of the kernels that changed since R3.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round_d.sh TAG
TAG=${1:-R4}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1
NOSEC="--no-secondary --no-cpu-baseline --no-parity"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 $NOSEC > $O/${TAG}_ncu_launch_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_c5_launches.csv \
    python tools/bench_configs.py c5 --nsteps 8192 > $O/${TAG}_ncu_launch_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k2_tet10' -s 6 -c 2 \
    -o $O/${TAG}_tet10 -f python tools/bench_configs.py c3 --curved surface --steps 2 > $O/${TAG}_ncu_tet10.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k2_hex20_steplane' -s 3 -c 1 \
    -o $O/${TAG}_hex20 -f python tools/bench_configs.py hex20 --steps 2 > $O/${TAG}_ncu_hex20.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_stream_kernel|gage_post_kernel' -s 8 -c 4 \
    -o $O/${TAG}_k3 -f python tools/bench_configs.py c5 --nsteps 4096 > $O/${TAG}_ncu_k3.log 2>&1
timeout 900 python tools/bench_configs.py c1 c1cli hex20 thick tri coat > $O/${TAG}_bench_configs.json 2> $O/${TAG}_bench_configs.err
tail -3 $O/${TAG}_pytest.log; cut -c1-600 $O/${TAG}_bench.json
