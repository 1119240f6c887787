for d in 2 3; do
echo "== FAV_SG_DBG=$d"
FAV_SG_DBG=$d timeout 200 python bench.py --steps 2 --warmup 1 --no-graph 2>&1 | grep "stem grad prof" | tail -1
done
