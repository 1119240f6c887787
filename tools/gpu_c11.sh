mkdir -p gpurun_out
FAV_TAP_PROF=1 FAV_BRANCH_STREAMS=0 timeout 300 python tools/tap_prof_arch.py r2plus1d_18 2> gpurun_out/c11_tap_prof_r21.txt > /dev/null; echo "exit $?"
awk '/profiled step/{p=1} p' gpurun_out/c11_tap_prof_r21.txt > /dev/null
python - <<'PY'
import re
lines=open("gpurun_out/c11_tap_prof_r21.txt").read().split("\n")
# second half = profiled step (the tool prints "=== profiled step" to stdout, so split by count)
recs=[l for l in lines if "tap prof" in l]
recs=recs[len(recs)//2:]
tot=0
out=[]
for l in recs:
    m=re.search(r"M (\S+) k(\S+) cin=(\d+) nkb=(\d+) bn=(\S+) stages=(\d+) grid=(\d+): per-CTA kclk total ([\d.]+), wait tempty ([\d.]+), full ([\d.]+), tiles ([\d.]+)", l)
    if m:
        out.append((float(m.group(8)), l[6:]))
        tot+=float(m.group(8))
print("launches", len(out), "sum kclk", tot)
for t,l in sorted(out, reverse=True)[:25]: print(l[:170])
PY
