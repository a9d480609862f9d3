CVB_WGRAD_TH=8 timeout 200 python -m pytest tests/test_ops_gpu.py -x -q -k "wgrad" > gpurun_out/r02z2_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r02z2_pytest.log
timeout 200 python -m pytest tests/test_ops_gpu.py -x -q -k "wgrad" > gpurun_out/r02z3_pytest.log 2>&1; echo rc=$?; tail -2 gpurun_out/r02z3_pytest.log
for s in "16 45 60 1024 512" "16 45 60 512 512" "16 90 120 256 256" "16 22 30 1024 1024" "16 180 240 256 128" "16 180 240 128 128"; do
  for t in 0 8 12; do CVB_WGRAD_TH=$t timeout 100 python tools/bench_wgrad.py $s 7 2>&1 | tail -1 | sed "s/^/th=$t  /"; done
done > gpurun_out/r02z_wgrad_th.txt 2>&1
cat gpurun_out/r02z_wgrad_th.txt
