timeout 600 python -m pytest tests/test_models_gpu.py -x -q -k "input_channel" > gpurun_out/r02v_pytest.log 2>&1; tail -15 gpurun_out/r02v_pytest.log
