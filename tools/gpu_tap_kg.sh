# per-tap conv kernel: k-blocks per barrier hand-off (FAV_TAP_KG; unset = planner's choice) — parity tests, then the step
# time and conv_tap family time of c2 (I3D) and c4 (r2plus1d_18) for each setting on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_i3d.py tests/test_gpu_resnet.py -m gpu -x -q --timeout 300 -p no:cacheprovider > gpurun_out/kg_tests.log 2>&1
echo "tests exit $?"; tail -2 gpurun_out/kg_tests.log
for k in ${KGS:-1 2 3 4 auto}; do
  for c in c2 c4; do
    echo -n "KG=$k $c: "
    if [ "$k" = auto ]; then unset FAV_TAP_KG; else export FAV_TAP_KG=$k; fi
    timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'conv_tap', round(d['kernels']['conv_tap']['ms_per_step'],3))"
  done
done
