#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, ncu launch list + full captures of the hot kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh TAG [quick]
TAG=${1:-R2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1
NOSEC="--no-secondary --no-cpu-baseline --no-parity"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 $NOSEC > $O/${TAG}_ncu_launch_bench.log 2>&1
# full captures: K1 + the quad kernel of the headline; TET10 kernels of config 3; the record kernel; K3
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k1_expand|k2_quad_planar|k2_quad_flat|k2_shell_vm' -s 6 -c 4 \
    -o $O/${TAG}_k1_k2 -f python bench.py --steps 2 --warmup 3 $NOSEC > $O/${TAG}_ncu_k1_k2.log 2>&1
if [ "$2" != "quick" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_tet10' -s 6 -c 2 \
    -o $O/${TAG}_tet10 -f python tools/bench_configs.py c3 --curved surface --steps 2 > $O/${TAG}_ncu_tet10.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k2_hex20_steplane' -s 3 -c 1 \
    -o $O/${TAG}_hex20 -f python tools/bench_configs.py hex20 --steps 2 > $O/${TAG}_ncu_hex20.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'record_points_dmma' -s 2 -c 2 \
    -o $O/${TAG}_record -f python tools/bench_record.py --only all --steps 64 > $O/${TAG}_ncu_record.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k3_stream_kernel|gage_post_kernel' -s 8 -c 4 \
    -o $O/${TAG}_k3 -f python tools/bench_configs.py c5 --nsteps 4096 > $O/${TAG}_ncu_k3.log 2>&1
timeout 600 python tools/bench_record.py > $O/${TAG}_record.json 2> $O/${TAG}_record.err
timeout 900 python tools/bench_cli.py --nx 500 --ny 500 --steps 2000 --shm > $O/${TAG}_cli.json 2> $O/${TAG}_cli.err
timeout 900 python tools/bench_configs.py c1 c1cli hex20 thick tri coat > $O/${TAG}_bench_configs.json 2> $O/${TAG}_bench_configs.err
fi
tail -3 $O/${TAG}_pytest.log; cut -c1-600 $O/${TAG}_bench.json
