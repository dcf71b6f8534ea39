#!/bin/bash
# chain queues (a second kernel for the paths inside clear media): images and ray counts against the single-kernel path, then throughput
B=./mu-lambda-raytracer_b200/rt_main
for w in final_scene cornell_smoke; do
  A="--world=$w --seed=42 --image_width=200 --samples_per_pixel=100"
  RT_PS_CHAINQ=0 timeout 120 $B $A --stats > /tmp/off.ppm 2>/tmp/off.err; echo "off rc=$?"
  RT_PS_CHAINQ_VERBOSE=1 timeout 120 $B $A --stats > /tmp/on.ppm 2>/tmp/on.err; echo "on rc=$?"; grep "chain queues" /tmp/on.err | head -12
  RT_PS_CHAINQ_LOG2_PATHS=16 RT_PS_CHAINQ_LOG2_CAP=10 timeout 120 $B $A --stats > /tmp/on2.ppm 2>/tmp/on2.err; echo "on2 (small launches, tiny queues) rc=$?"
  python - "$w" <<'PY'
import sys, json, numpy as np
def load(p):
    t = open(p).read().split()
    return np.array(t[4:], dtype=np.int64)
def st(p):
    return json.loads(open(p).read().strip().splitlines()[-1])
a, b, c = load("/tmp/off.ppm"), load("/tmp/on.ppm"), load("/tmp/on2.ppm")
ra, rb, rc = st("/tmp/off.err"), st("/tmp/on.err"), st("/tmp/on2.err")
print(sys.argv[1], "values", a.size, "| queues: differ", int((a != b).sum()), "max", int(np.abs(a - b).max()), "| small queues: differ", int((a != c).sum()), "max", int(np.abs(a - c).max()),
      "| rays", ra["rays"], rb["rays"], rc["rays"], "| launches", ra["kernel_launches"], rb["kernel_launches"], rc["kernel_launches"])
PY
done
C4="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=1000"
C3="--world=cornell_smoke --seed=42 --aspect_ratio=1:1 --image_width=600 --samples_per_pixel=1000"
run() { cfg=$1; shift; for rep in 1 2; do env "$@" timeout 120 $B $cfg --stats 2>&1 >/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['mpaths_per_s'], round(d['rays']/d['paths'],4), d['kernel_launches'])"; done; }
echo "== C4"; run "$C4" RT_PS_CHAINQ=0; run "$C4" RT_PS_CHAINQ=1; run "$C4" RT_PS_CHAINQ_LOG2_PATHS=26; run "$C4" RT_PS_CHAINQ_LOG2_PATHS=29 RT_PS_CHAINQ_LOG2_CAP=26
echo "== C3"; run "$C3" RT_PS_CHAINQ=0; run "$C3" RT_PS_CHAINQ=1
