timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "bilinear" > gpurun_out/r02u_pytest.log 2>&1; tail -2 gpurun_out/r02u_pytest.log
for s in "16 180 240 128" "16 45 60 512"; do timeout 100 python tools/bench_bn.py $s 2>&1 | grep -E "shape|bilinear"; done
for st in 1 0 1 0; do timeout 100 python tools/bench_conv.py 16 360 480 64 64 $st 2>&1 | tail -1; done
timeout 100 python tools/bench_conv.py 16 360 480 128 64 1 2>&1 | tail -1
timeout 100 python tools/bench_conv.py 16 360 480 128 64 0 2>&1 | tail -1
