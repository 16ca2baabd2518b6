N=${1:-4}
nvidia-smi topo -m 2>&1 | head -20
nproc; cat /sys/fs/cgroup/cpu.max; numactl -H 2>/dev/null | head -8 || lscpu | grep -i numa
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -c 300 gpurun_out/r02_bench_n$N.err
python - <<EOF
import json
d=json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value",d["value"],d["ms_per_step"],"e2e",d["e2e"]["value"],"per-rank h2d",d["e2e"]["h2d_gbs_per_rank"],d["e2e"]["host_placement"])
print(d["gather"]); print("K1a frac",d["roofline"]["frac"], "clocks", d["clocks"])
EOF
