# A/B of two builds of libfav.so on one box: the in-tree build vs flickering_adversarial_video_b200/libfav_ab.so (a variant built
# with an extra -D switch in a scratch directory); c2 / c4 step times, alternating, two rounds
P=flickering_adversarial_video_b200
cp $P/libfav.so /tmp/libfav_base.so
run() { timeout 300 python bench.py --config $1 --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items() if v['ms_per_step']>0.3})"; }
for r in 1 2; do
  for v in base ab; do
    if [ $v = ab ]; then cp $P/libfav_ab.so $P/libfav.so; else cp /tmp/libfav_base.so $P/libfav.so; fi
    for c in ${CONFIGS:-c2 c4}; do echo -n "$v $c: "; run $c; done
    for a in $ARCHS; do echo -n "$v "; timeout 300 python tools/bench_arch.py $a 2>&1 | tail -1 | cut -c1-140; done
  done
done
cp /tmp/libfav_base.so $P/libfav.so
