# usage: bash tools/ab_env.sh VAR v1 v2 ... : I3D bench step time for each value of an environment switch (same box)
var=$1; shift
for v in "$@"; do
  echo -n "$var=$v  "
  env $var=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
done
