mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/i3d_parity.log
if [ "$1" != "quick" ]; then
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -s --timeout 120 -p no:cacheprovider > gpurun_out/ops.log 2>&1; echo "ops exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/ops.log
fi
timeout 900 python -m pytest tests/test_gpu_i3d.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/i3d.log 2>&1; echo "i3d exit $?" >> gpurun_out/summary.txt
tail -8 gpurun_out/i3d.log; cat gpurun_out/summary.txt
