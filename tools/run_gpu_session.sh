set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/r02_bench_under_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_umma -s 60 -c 60 --csv --log-file gpurun_out/r02_conv_dram.csv python tools/profile_forward.py 3 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:conv_umma -s 60 -c 40 -o /tmp/r02_conv_full python tools/profile_forward.py 2 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02_conv_full.ncu-rep > gpurun_out/r02_ncu_conv_forward.txt
ncu --set full --clock-control none -k regex:"attention|down2|up2|pack_input" -s 9 -c 9 -o /tmp/r02_other_full python tools/profile_forward.py 2 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02_other_full.ncu-rep > gpurun_out/r02_ncu_other.txt
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 62 -c 2 -o gpurun_out/r02_conv_l0_src python tools/profile_forward.py 2 > /dev/null 2>&1
ls -la gpurun_out/r02_*; du -sh gpurun_out
