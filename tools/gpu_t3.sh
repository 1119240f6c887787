# (3,1,1) temporal-sharing conv (conv_t3.cu; FAV_T3=0 keeps the per-tap kernel): torch-stack parity tests, then the
# r2plus1d_18 step (c4) under both settings and the t3 launches' wait breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_torch_api.py tests/test_gpu_baseline_shapes.py -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/t3_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/t3_tests.log
for v in 0 1; do
  echo -n "FAV_T3=$v c4: "
  FAV_T3=$v timeout 300 python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items() if v['ms_per_step']>0.05})"
done
FAV_TAP_PROF=1 timeout 300 python tools/tap_prof_arch.py r2plus1d_18 2>&1 | grep "t3 prof" | sort | uniq -c
