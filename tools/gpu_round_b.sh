TAG=R2b
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_tet10' -s 4 -c 4 \
    -o $O/${TAG}_tet10 -f python tools/bench_configs.py c3 --curved surface --steps 2 > $O/${TAG}_ncu_tet10.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'record_points_dmma' -s 2 -c 2 \
    -o $O/${TAG}_record -f python tools/bench_record.py --only all --steps 64 > $O/${TAG}_ncu_record.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k3_stream_kernel|gage_post_kernel' -s 8 -c 4 \
    -o $O/${TAG}_k3 -f python tools/bench_configs.py c5 --nsteps 4096 > $O/${TAG}_ncu_k3.log 2>&1
timeout 600 python tools/bench_record.py > $O/${TAG}_record.json 2> $O/${TAG}_record.err
timeout 900 python tools/bench_cli.py --nx 500 --ny 500 --steps 2000 --shm > $O/${TAG}_cli.json 2> $O/${TAG}_cli.err
timeout 900 python tools/bench_configs.py c1 c1cli hex20 thick tri coat > $O/${TAG}_bench_configs.json 2> $O/${TAG}_bench_configs.err
ls -la $O | tail -15
