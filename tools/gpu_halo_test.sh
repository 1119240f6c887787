mkdir -p gpurun_out
for mode in 1 0; do
  echo "=== FAV_SWZ_BASE_OFFSET=$mode ==="
  FAV_SWZ_BASE_OFFSET=$mode timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 -p no:cacheprovider -k "conv3d" 2>&1 | tail -12
done
