#!/bin/bash
# per-kernel durations of one CLI render (ncu launch list): tools/launch_list.sh <world> <width> <spp> [ENV=...]
B=./mu-lambda-raytracer_b200/rt_main
W=$1; IW=$2; SPP=$3; shift 3
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file /tmp/launches.csv $B --world=$W --seed=42 --aspect_ratio=1:1 --image_width=$IW --samples_per_pixel=$SPP > /dev/null 2>/tmp/ll.err
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('/tmp/launches.csv')) if len(r)>10 and r[0].isdigit()]
from collections import defaultdict
t=defaultdict(float); n=defaultdict(int); seq=[]
for r in rows:
    k=r[4].split('(')[0].replace('void rtb::','')[:60]; t[k]+=float(r[-1]); n[k]+=1; seq.append((k,float(r[-1])))
tot=sum(t.values())
for k,v in sorted(t.items(), key=lambda kv:-kv[1])[:8]: print("%10.3f ms %6.2f%% %4d  %s"%(v/1e6,100*v/tot,n[k],k))
print("first launches:", [(k[:14], round(v/1e6,3)) for k,v in seq[:14]])
PY
