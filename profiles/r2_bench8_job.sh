for wl in encode256 roundtrip512 slide; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload $wl --no-extras > gpurun_out/r2h_bench8_$wl.json 2> gpurun_out/r2h_bench8_$wl.err
  cut -c1-200 gpurun_out/r2h_bench8_$wl.json
done
