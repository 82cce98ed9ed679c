#!/bin/bash
# 2-GPU visit: sharded tests, both arms of the driver's torchrun command at N = 2, config 4 on two GPUs.
# Usage: gpurun --gpus 2 -- bash tools/gpu_round_n2.sh TAG
TAG=${1:-R4n2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $O/${TAG}_pytest_sharded.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_sharded.log
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
timeout 600 $TR bench.py --gpus 2 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 500 $TR tools/bench_c4.py > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err
tail -2 $O/${TAG}_pytest_sharded.log; cut -c1-300 $O/${TAG}_bench.json; echo; cut -c1-300 $O/${TAG}_bench_c4.json; tail -3 $O/${TAG}_bench.err
