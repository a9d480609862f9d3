set -x
for s in "16 45 60 1024 512" "16 45 60 512 512" "16 90 120 256 256" "16 22 30 1024 1024" "16 180 240 256 128"; do
  tag=$(echo $s | tr ' ' '_')
  ncu --set full --clock-control none -k regex:"conv_wgrad_kernel" --launch-skip 1 -c 1 -o gpurun_out/r02p_wgrad_$tag python tools/bench_wgrad.py $s 3 > gpurun_out/r02p_ncu_$tag.log 2>&1
  tail -1 gpurun_out/r02p_ncu_$tag.log
done
