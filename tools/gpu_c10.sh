mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf -x > gpurun_out/c10_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/c10_pytest.log
for cfg in c2 c1 c4; do for v in 0 1; do
  echo -n "$cfg FAV_STEM_RAW=$v: "
  FAV_STEM_RAW=$v timeout 300 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done; done
