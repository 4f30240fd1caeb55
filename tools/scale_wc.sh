#!/bin/bash
# e2e at N GPUs: write-combined input buffer and pipe depth variants (run under gpurun --gpus N)
N=${1:-8}
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-profile-pass --no-other-configs 2>> gpurun_out/scale_wc.err | grep '^{' | sed "s/^{/{\"variant\": \"$name\", /" >> gpurun_out/scale_wc.jsonl
}
run local_wc DLKA_HOST_NUMA=local DLKA_HOST_WC=1
run local_wc_depth3 DLKA_HOST_NUMA=local DLKA_HOST_WC=1 DLKA_PIPE_DEPTH=3
run local_depth3 DLKA_HOST_NUMA=local DLKA_PIPE_DEPTH=3
python - <<'PY'
import json
for l in open("gpurun_out/scale_wc.jsonl"):
    d = json.loads(l)
    print(d["variant"], d["n_gpus"], "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4))
PY
