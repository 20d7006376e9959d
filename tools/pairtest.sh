R2DM_PAIR=2 timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -x -q -k "test_conv or test_fused_groupnorm_conv or forward_matches" 2>&1 | tail -4
R2DM_PAIR=2 R2DM_PRINT_PROFILE=1 timeout 200 python tools/profile_forward.py 3 > gpurun_out/s26_profile.txt 2>&1; head -3 gpurun_out/s26_profile.txt
