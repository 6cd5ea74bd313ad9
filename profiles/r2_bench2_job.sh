for wl in encode256 roundtrip512 slide; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload $wl --no-extras > gpurun_out/r2d_bench2_$wl.json 2> gpurun_out/r2d_bench2_$wl.err
  tail -c 600 gpurun_out/r2d_bench2_$wl.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2d_bench2_$wl.json").read().strip().splitlines()[-1])
print("$wl", d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"], d["detail"].get("host_numa_node_rank0"), d["clocks"])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | tail -c 700
