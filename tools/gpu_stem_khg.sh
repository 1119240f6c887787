# temporal-sharing stem: kh taps per weight block (FAV_STEM_TS_KHG) — parity tests at the default, then stem_conv time per value
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_i3d.py tests/test_gpu_resnet.py -m gpu -x -q --timeout 300 -p no:cacheprovider > gpurun_out/ts_tests.log 2>&1
echo "tests exit $?"; tail -2 gpurun_out/ts_tests.log
for k in ${KHGS:-1 2 3}; do
  echo -n "KHG=$k: "
  FAV_STEM_TS_KHG=$k timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels']['stem_conv']['ms_per_step'])"
  FAV_STEM_TS_KHG=$k FAV_STEM_PROF=1 timeout 300 python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline --sustained-sec 0 2>&1 | grep "stem ts prof" | tail -1
done
FAV_STEM_TS_KHG=3 timeout 300 python -m pytest tests/test_gpu_i3d.py -m gpu -x -q --timeout 300 -p no:cacheprovider -k "layer or forward" 2>&1 | tail -1
