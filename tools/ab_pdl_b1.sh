for p in 0 1 0 1; do
  echo -n "PDL=$p "
  FAV_PDL=$p timeout 300 python bench.py --frames 90 --batch 1 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
done
