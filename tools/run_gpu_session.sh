python -m pytest tests -m gpu -q -x > gpurun_out/r2_s16_pytest.txt 2>&1; tail -3 gpurun_out/r2_s16_pytest.txt
for i in 1 2 3; do python tools/dbg_fold2.py fp32 8 H 6 2>&1 | tail -1; done
python bench.py --quick > gpurun_out/r2_s16_bench.json 2> gpurun_out/r2_s16_bench.err; cut -c1-200 gpurun_out/r2_s16_bench.json
