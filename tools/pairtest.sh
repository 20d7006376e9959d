for shape in "256 256 16 256 1" "512 512 8 128 1" "128 256 32 512 0" "256 256 8 128 1"; do
  for v in 0 1; do
    echo "== shape $shape PAIR=$v"
    R2DM_PAIR=$v ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_umma|conv_pair" -s 2 -c 3 python tools/ncu_conv.py $shape 5 2>&1 | grep -E "gpu__time_duration" | awk '{print $NF, $(NF-1)}' | tr '\n' ' '; echo
  done
done
