mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/c4_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/c4_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench exit $?"; tail -3 gpurun_out/c4_bench.err
for c in c1 c3 c4 c5 c5m; do timeout 600 python bench.py --config $c --steps 20 --warmup 3 --sustained-sec 0 > gpurun_out/c4_bench_$c.json 2> gpurun_out/c4_bench_$c.err; echo "bench $c exit $?"; tail -2 gpurun_out/c4_bench_$c.err; done
echo -n "pair na=3: "; FAV_HALO_PAIR_NA=3 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels']['conv_halo']['ms_per_step'])"
timeout 600 python tools/bench_arch.py > gpurun_out/c4_bench_arch.log 2>&1
for a in "r3d_18 16 16" "r2plus1d_18 16 16"; do n=$(echo $a | tr ' ' '_'); timeout 600 python tests/gpu_torch_reference_timing.py $a > gpurun_out/c4_torch_ref_$n.log 2>&1; timeout 600 python tests/gpu_torch_reference_timing.py $a --tf32 > gpurun_out/c4_torch_ref_tf32_$n.log 2>&1; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c4_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["config"]["name"], round(d["ms_per_step"], 3), "ms", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "step frac", round(d["roofline"]["step"]["frac"], 3), "sust", d.get("sustained"))
        print("   ", {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 2 gpurun_out/c4_torch_ref_*.log; cat gpurun_out/c4_bench_arch.log | tail -3
