mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_ops.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -s > gpurun_out/c5_pytest_eval.log 2>&1; echo "pytest eval exit $?"
tail -15 gpurun_out/c5_pytest_eval.log
grep -E "device bytes|vs oracle" gpurun_out/c5_pytest_eval.log
