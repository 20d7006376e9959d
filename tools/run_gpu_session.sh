python -m pytest tests -m gpu -q > gpurun_out/r2_s30_pytest.txt 2>&1; tail -2 gpurun_out/r2_s30_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2_s30_bench.json 2> gpurun_out/r2_s30_bench.err; tail -2 gpurun_out/r2_s30_bench.err; cut -c1-160 gpurun_out/r2_s30_bench.json
