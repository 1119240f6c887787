# Multi-GPU checks on one box (gpurun --gpus N): bench at N ranks for c2 and c3 (clean teardown expected: exit 0 without
# the watchdog), the sharded per-pixel attack's exchange, and the 1-GPU numbers of the same box for the efficiency ratio.
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/multi_$name.log 2> gpurun_out/multi_$name.err; echo "$name exit $?"; tail -c 600 gpurun_out/multi_$name.err | tail -3; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run c2_n1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0
run c2_n$N $TR bench.py --gpus $N --steps 20 --warmup 3 --sustained-sec 0
run c3_n1 python bench.py --config c3 --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0
run c3_n$N $TR bench.py --gpus $N --config c3 --steps 20 --warmup 3 --sustained-sec 0
run c4_n1 python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0
run c4_n$N $TR bench.py --gpus $N --config c4 --steps 20 --warmup 3 --sustained-sec 0
run sparse_i3d_n1 python tools/bench_sparse_exchange.py i3d 2 64
run sparse_i3d_n$N $TR tools/bench_sparse_exchange.py i3d 2 64
run sparse_r3d_n$N $TR tools/bench_sparse_exchange.py r3d_18 4 16
for f in gpurun_out/multi_*.log; do echo "== $f"; python - "$f" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d = json.loads(ln)
        keep = {k: d[k] for k in ("n_gpus", "ms_per_step", "value", "exchange_ms", "exchange_bytes", "clip_frames_per_sec") if k in d}
        if "e2e" in d: keep["e2e"] = d["e2e"]["value"]
        print(keep)
PY
done
