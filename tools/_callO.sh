timeout 60 tools/exp/exp_mma_rate2 > gpurun_out/r02y_mma_rate2.txt 2>&1; echo rc=$?; cat gpurun_out/r02y_mma_rate2.txt
