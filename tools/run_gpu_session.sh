python -m pytest tests -m gpu -q > gpurun_out/r2_s26_pytest.txt 2>&1; tail -3 gpurun_out/r2_s26_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2_s26_bench.json 2> gpurun_out/r2_s26_bench.err; tail -2 gpurun_out/r2_s26_bench.err; cut -c1-200 gpurun_out/r2_s26_bench.json
python bench.py --impl reference --steps 3 --warmup 3 2>/dev/null | cut -c1-300
