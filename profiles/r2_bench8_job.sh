nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1
for wl in encode256; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload $wl --no-extras > gpurun_out/r2g_bench8_$wl.json 2> gpurun_out/r2g_bench8_$wl.err
  cut -c1-260 gpurun_out/r2g_bench8_$wl.json
done
python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_bench8_encode256.json").read().strip().splitlines()[-1])
print(round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["detail"]["host_numa_node_rank0"])
PY
head -12 gpurun_out/r2f_topo.txt | cut -c1-200
