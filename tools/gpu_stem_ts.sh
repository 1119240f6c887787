# A/B of the temporal-sharing stem kernel (FAV_STEM_TS=0 keeps the raw-row kernel): layer-wise parity tests first, then
# the I3D bench step and the r3d_18 step under both settings on the same box.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_i3d.py tests/test_gpu_resnet.py -m gpu -x -q --timeout 300 -p no:cacheprovider > gpurun_out/ts_tests.log 2>&1
echo "tests exit $?"; tail -6 gpurun_out/ts_tests.log
for v in 0 1; do
  echo -n "FAV_STEM_TS=$v  c2: "
  FAV_STEM_TS=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>gpurun_out/ts_bench_$v.err | tail -1 > gpurun_out/ts_bench_$v.json
  python -c "import json; d=json.load(open('gpurun_out/ts_bench_$v.json')); print(d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()} if 'kernels' in d else '')"
done
for v in 0 1; do
  echo "FAV_STEM_TS=$v r3d_18:"; FAV_STEM_TS=$v timeout 300 python tools/bench_arch.py r3d_18 2>&1 | tail -3
done
