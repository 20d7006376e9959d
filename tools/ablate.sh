for v in "R2DM_CONV_DEBUG=7 R2DM_XF_DEBUG=1" "R2DM_CONV_DEBUG=3 R2DM_XF_DEBUG=1" "R2DM_CONV_DEBUG=5 R2DM_XF_DEBUG=1"; do
  echo "== $v"; env $v R2DM_PRINT_PROFILE=1 timeout 100 python tools/profile_forward.py 2 2>&1 | grep -E "^  # *(2|3|8|10|16|18|26|38|50|59) |^conv3x3"
done
