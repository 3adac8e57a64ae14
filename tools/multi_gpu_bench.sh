#!/bin/bash
# N-GPU measurements the driver's scaling run does not cover: camera-sharded latency
# mode (configs 2 and 4) and the R101 400x400x16 stress config, next to the replica mode.
#   gpurun --gpus N -- bash tools/multi_gpu_bench.sh N
N=${1:-2}
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 python -m torch.distributed.run --nnodes=1 \
      --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
      > gpurun_out/mg_${name}_n$N.json 2> gpurun_out/mg_${name}_n$N.err
  echo "$name exit $?"; grep -c "NCCL INFO" gpurun_out/mg_${name}_n$N.err
  grep -h "comm .* rank .* nranks" gpurun_out/mg_${name}_n$N.err | head -2
}
run replicas --steps 30 --warmup 5 --no-cpu-baseline --no-extras --no-profile
run shard_finetune --shard camera --steps 30 --warmup 5 --no-cpu-baseline --no-extras --no-profile
run shard_traj --shard camera --config traj --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-profile
run stress --config stress --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-profile
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/mg_*_n$N.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 3), "ms/step e2e", d.get("e2e", {}).get("value"))
PY
