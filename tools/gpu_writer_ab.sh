#!/bin/bash
# A/B of the results-database file stage: pwritev helper threads vs copies into a shared mapping of the file (FSR_RDB_MMAP=1)
TAG=${1:-R4wr}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_rdb.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest.log
timeout 600 python tools/bench_cli.py --nx 500 --ny 500 --steps 2000 --shm --writer-ab --skip-all > $O/${TAG}_cli.json 2> $O/${TAG}_cli.err
tail -3 $O/${TAG}_pytest.log; tail -3 $O/${TAG}_cli.err
python - $O/${TAG}_cli.json <<'P'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); s=d["split"]
    print(d["results_database"], d["writer"], "wall %.2f s  loop %.2f  file %.2f  d2h %.2f  dev %.3f  ctx %.2f" % (d["seconds_wall"], s.get("time_loop_s",0), s.get("file_s",0), s.get("d2h_s",0), s.get("device_s",0), s.get("cuda_context_s",0)), d.get("parity_max_rel_vs_oracle_float_file"))
P
