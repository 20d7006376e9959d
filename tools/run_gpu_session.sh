python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_boundary_gpu.py -m gpu -q -x > gpurun_out/r2_s29_pytest.txt 2>&1; tail -2 gpurun_out/r2_s29_pytest.txt
( python tools/fwd_time.py bf16 8
  R2DM_LIB_PATH=$PWD/r2dm_b200/libr2dm_b200_alt.so python tools/fwd_time.py bf16 8
  python tools/fwd_time.py bf16 8
  R2DM_LIB_PATH=$PWD/r2dm_b200/libr2dm_b200_alt.so python tools/fwd_time.py bf16 8 ) > gpurun_out/r2_s29_fwd.txt 2>&1
grep FWD gpurun_out/r2_s29_fwd.txt | cut -c1-200
