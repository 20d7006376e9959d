for v in "A=0" "R2DM_XF_DEBUG=1" "R2DM_XF_DEBUG=2" "R2DM_XF_DEBUG=3" "R2DM_CONV_DEBUG=1" "R2DM_CONV_DEBUG=4" "R2DM_PDL=0"; do
  echo "== $v"; env $v R2DM_PRINT_PROFILE=1 timeout 100 python tools/profile_forward.py 2 2>&1 | grep -E "^  # *(2|3|8|10|11|16|18|19|24|26|27|38|50|65|72) |^conv3x3"
done
