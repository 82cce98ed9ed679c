#!/bin/bash
# 8-GPU visit: sharded tests, headline bench (weak C2 + strong C3), config 4.  Usage: gpurun --gpus 8 -- bash tools/gpu_round_n8.sh TAG
TAG=${1:-R2n8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $O/${TAG}_pytest_sharded.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_sharded.log
timeout 900 $TR bench.py --gpus 8 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err
timeout 900 $TR tools/bench_c4.py > $O/${TAG}_bench_c4_n8.json 2> $O/${TAG}_bench_c4_n8.err
tail -2 $O/${TAG}_pytest_sharded.log; cut -c1-300 $O/${TAG}_bench_n8.json; echo; cut -c1-300 $O/${TAG}_bench_c4_n8.json
