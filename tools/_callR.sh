export CVB_WGRAD_CLUSTER=2
timeout 120 python -m pytest tests/test_ops_gpu.py -x -q -k "wgrad" > gpurun_out/r02z4_pytest.log 2>&1; echo rc=$?; tail -5 gpurun_out/r02z4_pytest.log
for s in "16 45 60 1024 512" "16 45 60 512 512" "16 90 120 256 256" "16 22 30 1024 1024" "16 180 240 256 128"; do
  for c in 0 2; do CVB_WGRAD_CLUSTER=$c timeout 60 python tools/bench_wgrad.py $s 7 2>&1 | tail -1 | sed "s/^/cluster=$c  /"; done
done > gpurun_out/r02z_wgrad_pair.txt 2>&1
cat gpurun_out/r02z_wgrad_pair.txt
