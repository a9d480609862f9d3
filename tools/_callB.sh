set -x
python -m pytest tests/test_ops_gpu.py -x -q -k "last_block or layout or bilinear" > gpurun_out/r02n_pytest_ops.log 2>&1; tail -3 gpurun_out/r02n_pytest_ops.log
python -m pytest tests/test_models_gpu.py -x -q -k "train_step_matches_oracle or eval or optimizer" > gpurun_out/r02n_pytest_models.log 2>&1; tail -3 gpurun_out/r02n_pytest_models.log
ncu --set full --clock-control none --import-source on -k regex:"bilinear2x_bwd|pool_bwd_bn_reduce|bilinear2x_fwd" --launch-skip 2 -c 3 -o gpurun_out/r02n_ew_full python tools/bench_bn.py 16 180 240 128 > gpurun_out/r02n_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bilinear2x_bwd" --launch-skip 2 -c 1 -o gpurun_out/r02n_bil_bwd_full python tools/bench_bn.py 16 180 240 128 > gpurun_out/r02n_ncu2.log 2>&1
