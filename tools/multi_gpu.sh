#!/bin/bash
# multi-GPU visit (gpurun --gpus N): sharded-render test, the strong-scaling bench (fixed C4 job) at every N up to the box
# size, the C5 job at the box size through bench.py --c5 and through the CLI
set -u
N=${1:-2}; TAG=${2:-r2}
timeout 600 python -m pytest tests/test_cli_and_multi.py -m gpu -x -q > gpurun_out/pytest_multi_$TAG.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/pytest_multi_$TAG.log
show() { python -c "import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); m=d.get('render_multi') or {}; print(round(d['value'],1), 'Mpaths/s; e2e', round(d['e2e']['value'],1), '; rt_render_multi', round(m.get('value',0),1), '; ms/step', round(d['ms_per_step'],1), '; checksum', d['image_checksum'])" 2>&1 | tail -1; }
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  out=gpurun_out/scale_${TAG}_n$n.json
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --cpu-spp 0 > $out 2> gpurun_out/scale_${TAG}_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 3 > $out 2> gpurun_out/scale_${TAG}_n$n.err
  fi
  echo "C4 strong n=$n rc=$? $(show $out)"
done
out=gpurun_out/scale_${TAG}_c5_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N --c5 --steps 2 --warmup 1 --e2e-steps 1 --multi-steps 1 > $out 2> gpurun_out/scale_${TAG}_c5_n$N.err
echo "C5 strong n=$N rc=$? $(show $out)"
echo "C5 cli n=$N $(./mu-lambda-raytracer_b200/rt_main --world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=3840 --samples_per_pixel=${C5_SPP:-4096} --gpus $N --stats 2>&1 >/dev/null | tail -1)"
