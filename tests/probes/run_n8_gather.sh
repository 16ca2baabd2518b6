#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
i=0
for extra in "--no-gather" "--gather-slots 4"; do
  i=$((i+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras $extra > gpurun_out/n8_g_$i.json 2> gpurun_out/n8_g_$i.err
  python - "$extra" $i <<'PY'
import json,sys
d=json.loads(open("gpurun_out/n8_g_%s.json"%sys.argv[2]).read().strip().splitlines()[-1])
g=d.get("gather") or {}
print(sys.argv[1],"value %.0f ms/step %.3f"%(d["value"],d["ms_per_step"]),"gather ms",g.get("ms_per_gather_rank0"),"verified",g.get("verified_all_ranks"), "power", d["clocks"].get("power_w_max"))
PY
done
