# halo convs: 9 taps per weight group for narrow tiles (FAV_HALO_BG9=0: at most 3) — op / network parity tests, then A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_i3d.py tests/test_gpu_resnet.py -m gpu -x -q --timeout 600 -p no:cacheprovider 2>&1 | tail -2
for r in 1 2; do for v in 0 1; do
  for c in c2 c4; do
    echo -n "BG9=$v $c: "
    FAV_HALO_BG9=$v timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'conv_halo', round(d['kernels']['conv_halo']['ms_per_step'],3))"
  done
done; done
