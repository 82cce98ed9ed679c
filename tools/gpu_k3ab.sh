O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "fatigue or gage or rainflow or k3 or coat or fpp" > $O/r4e_pytest.log 2>&1; tail -2 $O/r4e_pytest.log
timeout 600 python tools/bench_configs.py c5 > $O/r4e_c5_16.json 2> $O/r4e_c5.err
FSR_K3_CHUNK=32 timeout 600 python tools/bench_configs.py c5 > $O/r4e_c5_32.json 2>> $O/r4e_c5.err
cut -c1-420 $O/r4e_c5_16.json $O/r4e_c5_32.json
