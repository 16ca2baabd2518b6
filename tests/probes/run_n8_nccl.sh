#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
for ctas in default 8 4 2; do
  if [ $ctas = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/n8_nccl_$ctas.json 2> gpurun_out/n8_nccl_$ctas.err
  python - $ctas <<'PY'
import json,sys
d=json.loads(open("gpurun_out/n8_nccl_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
print("NCCL_MAX_CTAS",sys.argv[1],"value %.0f ms/step %.3f"%(d["value"],d["ms_per_step"]),"gather ms",d["gather"]["ms_per_gather_rank0"],"verified",d["gather"]["verified_all_ranks"])
PY
done
