set -x
python tools/hbm_rw.py > gpurun_out/r02m_hbm_rw.txt 2>&1
python tools/bench_bn.py 16 180 240 128 > gpurun_out/r02m_ew_180x240x128.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bilinear2x|pool_bwd_bn_reduce|maxpool_fwd|bn_reduce_kernel" -c 14 -o gpurun_out/r02m_ew_full python tools/bench_bn.py 16 180 240 128 > gpurun_out/r02m_ncu.log 2>&1
ls -la gpurun_out/
