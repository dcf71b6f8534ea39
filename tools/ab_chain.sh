#!/bin/bash
# chain-phase A/B (run under gpurun): identical images with the phase on and off, then C4 / C3 throughput over the thresholds
B=./mu-lambda-raytracer_b200/rt_main
C4="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=1000"
C3="--world=cornell_smoke --seed=42 --aspect_ratio=1:1 --image_width=600 --samples_per_pixel=1000"
C2="--world=random --seed=42 --aspect_ratio=3:2 --image_width=1200 --samples_per_pixel=500 --aperture=0.1 --focus_dist=10.0"
SMALL4="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=200 --samples_per_pixel=200"
SMALL3="--world=cornell_smoke --seed=42 --aspect_ratio=1:1 --image_width=200 --samples_per_pixel=200"
for S in "$SMALL4" "$SMALL3"; do
  RT_PS_CHAIN_TRIG=0 timeout 120 $B $S > /tmp/off.ppm 2>/dev/null; echo "off rc=$?"
  timeout 120 $B $S > /tmp/on.ppm 2>/dev/null; echo "on rc=$?"
  RT_PS_VARIANT=3 RT_PS_CHAIN_TRIG=8 RT_PS_CHAIN_MIN=1 timeout 120 $B $S > /tmp/on2.ppm 2>/dev/null; echo "on2 rc=$?"
  python - <<'PY'
import numpy as np
def load(p):
    t = open(p).read().split()
    return np.array(t[4:], dtype=np.int64)
a, b, c = load("/tmp/off.ppm"), load("/tmp/on.ppm"), load("/tmp/on2.ppm")
print("values", a.size, "on != off:", int((a != b).sum()), "max diff", int(np.abs(a - b).max()), "| on2 != off:", int((a != c).sum()), "max diff", int(np.abs(a - c).max()))
PY
done
run() { cfg=$1; shift; for rep in 1 2; do env "$@" timeout 120 $B $cfg --stats 2>&1 >/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['mpaths_per_s'], round(d['rays']/d['paths'],4))"; done; }
echo "== C4"
run "$C4" RT_PS_CHAIN_TRIG=0
run "$C4" RT_PS_CHAIN_TRIG=24
run "$C4" RT_PS_VARIANT=3 RT_PS_CHAIN_TRIG=0
for trig in 16 24 32 40 48; do for min in 8 12 16; do
  run "$C4" RT_PS_VARIANT=3 RT_PS_CHAIN_TRIG=$trig RT_PS_CHAIN_MIN=$min
done; done
echo "== C3"
run "$C3" RT_PS_CHAIN_TRIG=0
run "$C3" RT_PS_CHAIN_TRIG=24
run "$C3" RT_PS_VARIANT=3 RT_PS_CHAIN_TRIG=0
run "$C3" RT_PS_VARIANT=3 RT_PS_CHAIN_TRIG=24
run "$C3" RT_PS_VARIANT=3 RT_PS_CHAIN_TRIG=32 RT_PS_CHAIN_MIN=16
echo "== C2"
run "$C2" RT_PS_CHAIN_TRIG=0
run "$C2" RT_PS_VARIANT=3
echo "== stats"
for e in "RT_PS_CHAIN_TRIG=0" "RT_PS_CHAIN_TRIG=24" "RT_PS_VARIANT=3 RT_PS_CHAIN_TRIG=0" "RT_PS_VARIANT=3 RT_PS_CHAIN_TRIG=24" "RT_PS_VARIANT=3 RT_PS_CHAIN_TRIG=40"; do
echo $e; env $e RT_PS_STATS=1 timeout 120 $B --world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=64 2>&1 >/dev/null | grep "persist stats"
done
