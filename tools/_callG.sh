timeout 300 python tools/lag_ab.py unet 0 1 2 3 4 6 8 12 23 > gpurun_out/r02s_lag_ab_unet.txt 2>&1; cat gpurun_out/r02s_lag_ab_unet.txt | tail -12
timeout 200 python tools/lag_ab.py segnet 0 2 4 8 > gpurun_out/r02s_lag_ab_segnet.txt 2>&1; cat gpurun_out/r02s_lag_ab_segnet.txt | tail -6
