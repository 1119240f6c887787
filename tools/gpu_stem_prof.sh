# FAV_STEM_PROF wait breakdown of the temporal-sharing stem kernel: 1 = as shipped, 3 = without the epilogue's work,
# 5 = without the MMAs (timing only; the un-graphed bench step is just the vehicle)
mkdir -p gpurun_out
for prof in ${PROFS:-1 3 5}; do
  FAV_STEM_PROF=$prof timeout 300 python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline --sustained-sec 0 2>&1 | grep "stem ts prof" | tail -1
done
