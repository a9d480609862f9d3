timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "wgrad" > gpurun_out/r02r_pytest.log 2>&1; tail -3 gpurun_out/r02r_pytest.log
for s in "16 45 60 1024 512" "16 45 60 512 512" "16 22 30 1024 1024" "16 22 30 512 1024" "16 180 240 128 128" "16 360 480 64 64" "16 360 480 128 64" "16 90 120 256 256"; do
  timeout 120 python tools/bench_wgrad.py $s 5 2>&1 | tail -1
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --launch-skip 4 -c 2 -k regex:"wgrad" python tools/bench_wgrad.py $s 3 2>/dev/null | grep -E "wgrad" | grep -v "^wgrad" | awk -F'","' '{print "   ", substr($5,1,40), $(NF)}'
done > gpurun_out/r02r_wgrad_split.txt 2>&1
cat gpurun_out/r02r_wgrad_split.txt
