timeout 60 tools/exp/exp_mma_pair > gpurun_out/r02y_mma_pair.txt 2>&1; echo rc=$?; cat gpurun_out/r02y_mma_pair.txt
