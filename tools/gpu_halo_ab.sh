# Planner A/B of the halo conv on one box: per-launch cycle counts (FAV_HALO_PROF) of the dominant 3x3x3 shapes under
# forced N splits / M-tile counts / CTA pairs, then the whole-step time of the promising settings.
mkdir -p gpurun_out
for cfg in "" "FAV_HALO_NT=2" "FAV_HALO_NT=2 FAV_HALO_MT=2" "FAV_HALO_2CTA=2" "FAV_HALO_2CTA=2 FAV_HALO_MT=2" "FAV_HALO_NT=3" "FAV_HALO_NT=2 FAV_HALO_2CTA=0"; do
  echo "=== $cfg" >> gpurun_out/halo_ab.txt
  env FAV_HALO_PROF=1 $cfg timeout 300 python tools/halo_sweep.py 2c 3b.b1b 3c.b1b 3c.b2b 4b.b1b 4f.b1b 2>> gpurun_out/halo_ab.txt >/dev/null
done
for cfg in "" "FAV_HALO_NT=2" "FAV_HALO_2CTA=2" "FAV_HALO_2CTA=0"; do
  echo -n "step [$cfg] " >> gpurun_out/halo_ab.txt
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels']['conv_halo']['ms_per_step'])" >> gpurun_out/halo_ab.txt
done
cat gpurun_out/halo_ab.txt
