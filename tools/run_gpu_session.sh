python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q > gpurun_out/r2_s24_pytest.txt 2>&1; tail -12 gpurun_out/r2_s24_pytest.txt
