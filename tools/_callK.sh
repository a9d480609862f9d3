timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "bn_relu or last_block or pool" > gpurun_out/r02x_pytest.log 2>&1; tail -2 gpurun_out/r02x_pytest.log
timeout 300 python tools/order_ab.py unet > gpurun_out/r02x_order_ab_unet.txt 2>&1; tail -9 gpurun_out/r02x_order_ab_unet.txt
