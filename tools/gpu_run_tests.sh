mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/i3d_parity.log
for f in ops golden i3d api; do
  if [ "$1" = "quick" ] && [ "$f" = "ops" ]; then continue; fi
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/$f.log 2>&1; echo "$f exit $?" >> gpurun_out/summary.txt
  tail -4 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
