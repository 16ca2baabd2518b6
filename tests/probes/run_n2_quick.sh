python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -c 400 gpurun_out/r02_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode tracking --steps 50 --warmup 3 --no-cpu > gpurun_out/r02_track_n2.json 2> gpurun_out/r02_track_n2.err; tail -c 400 gpurun_out/r02_track_n2.err
python - <<EOF
import json
d=json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print("N2 value",d["value"],d["ms_per_step"],d["ms_per_step_per_rank"],"e2e",d["e2e"]["value"],d["e2e"]["h2d_gbs_per_rank"])
print(d["gather"]); print("K1a frac",d["roofline"]["frac"],d["clocks"])
t=json.loads(open("gpurun_out/r02_track_n2.json").read().strip().splitlines()[-1])
print("track N2",t["value"],t["ms_per_step"],t["gather"],t["clocks"],"e2e",t["e2e"]["value"])
