# Round-end verification on one box: the -m gpu suite, smoke(), the default bench line (with its CPU baseline and the
# sustained segment), the reference arm, every other BASELINE configuration.
mkdir -p gpurun_out
export TAG=${TAG:-close}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -rf > gpurun_out/${TAG:-close}_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/${TAG:-close}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" > gpurun_out/${TAG:-close}_smoke.log 2>&1; echo "smoke exit $?"; grep -E "smoke|SMOKE" gpurun_out/${TAG:-close}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG:-close}_bench.json 2> gpurun_out/${TAG:-close}_bench.err; echo "bench exit $?"; tail -2 gpurun_out/${TAG:-close}_bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG:-close}_bench_reference.json 2> gpurun_out/${TAG:-close}_bench_reference.err; echo "reference exit $?"; tail -c 400 gpurun_out/${TAG:-close}_bench_reference.json
for c in c1 c3 c4 c5 c5m; do timeout 600 python bench.py --config $c --steps 20 --warmup 3 --sustained-sec 0 --no-cpu-baseline > gpurun_out/${TAG:-close}_bench_$c.json 2> gpurun_out/${TAG:-close}_bench_$c.err; echo "bench $c exit $?"; done
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ.get("TAG", "close"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("impl") == "reference":
            print(f, "reference", d.get("value"), d.get("unit"), d.get("cpu_baseline", {}).get("cores"))
            continue
        print(f, d["config"].get("name"), round(d["ms_per_step"], 3), "ms", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "step", d["roofline"]["step"], "sustained", d.get("sustained"), "cpu", d.get("cpu_baseline", {}).get("value"), "traffic_err", d["roofline"].get("traffic_error"))
    except Exception as e:
        print(f, "ERR", e)
PY
