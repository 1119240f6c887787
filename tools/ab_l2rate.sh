for r in 70 100 55 70 100; do
  echo -n "L2RATE=$r "
  FAV_HALO_L2RATE=$r timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
done
