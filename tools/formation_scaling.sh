#!/bin/bash
# Strong scaling of one formation over the GPUs of a node: tools/formation_scaling.sh <world> [drone counts ...]
# (run once per world size under `gpurun --gpus <world>`; results land in gpurun_out/formation_multi_gpu_<world>_<drones>.json)
W=${1:-1}; shift
SIZES=${@:-"16384 65536 262144"}
for n in $SIZES; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      tools/formation_multi_gpu.py --drones $n --ticks 6 > gpurun_out/formation_scaling_w${W}_n${n}.log 2>&1
  echo "world $W drones $n rc=$?"
done
