python -m pytest tests -m gpu -q > gpurun_out/r2_s10_pytest.txt 2>&1; tail -8 gpurun_out/r2_s10_pytest.txt
python - <<'PY' 2>&1 | tail -5
import torch, json, sys
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.gpu_eager_baseline(torch.device('cuda', 0), 8), indent=1))
PY
