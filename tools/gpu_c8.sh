mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_i3d.py tests/test_gpu_resnet.py tests/test_gpu_eval.py -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/c8_pytest.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/c8_pytest.log
for cfg in c2 c4; do for v in 2 0; do
  echo -n "$cfg FAV_TAP_ACC=$v: "
  FAV_TAP_ACC=$v timeout 300 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done; done
FAV_TAP_PROF=1 FAV_BRANCH_STREAMS=0 timeout 300 python tools/profile_step.py 2> gpurun_out/c8_tap_prof.txt > /dev/null
grep "tap prof" gpurun_out/c8_tap_prof.txt | tail -38 | cut -c7-150 | head -8
