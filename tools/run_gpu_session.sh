( python tools/fwd_time.py bf16 8
  R2DM_OPT_FIRST_DELAY_NS=400 python tools/fwd_time.py bf16 8
  R2DM_OPT_FIRST_DELAY_NS=800 R2DM_OPT_FIRST_DELAY_STAGES=2 python tools/fwd_time.py bf16 8
  R2DM_OPT_FIRST_DELAY_NS=1500 R2DM_OPT_FIRST_DELAY_STAGES=2 python tools/fwd_time.py bf16 8
  python tools/fwd_time.py bf16 8 ) > gpurun_out/r2_s22_fwd.txt 2>&1
grep FWD gpurun_out/r2_s22_fwd.txt | cut -c1-200
