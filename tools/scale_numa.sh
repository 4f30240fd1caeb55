#!/bin/bash
# e2e scaling under the three host-placement policies (run under gpurun --gpus N): one line per (N, policy) in gpurun_out/scale_numa.jsonl
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
for pol in local off interleave; do
  for n in $N; do
    if [ "$n" = "1" ]; then
      DLKA_HOST_NUMA=$pol timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-profile-pass 2>> gpurun_out/scale_numa.err | sed "s/^{/{\"policy\": \"$pol\", /" >> gpurun_out/scale_numa.jsonl
    else
      DLKA_HOST_NUMA=$pol timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-profile-pass 2>> gpurun_out/scale_numa.err | grep '^{' | sed "s/^{/{\"policy\": \"$pol\", /" >> gpurun_out/scale_numa.jsonl
    fi
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/scale_numa.jsonl"):
    d = json.loads(l)
    print(d["policy"], d["n_gpus"], "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), d["config"]["host"][-60:])
PY
