#!/bin/bash
for cfg in "MPE_K1B_POOL=0" "MPE_K1B_POOL=4"; do
env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu --no-extras --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', d['value'], d['ms_per_step'], d['ms_per_step_per_rank'], 'gather', d['gather']['ms_per_gather_rank0'])"
done
