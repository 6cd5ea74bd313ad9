python -m pytest tests/test_gpu_graphs.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -15
python bench.py --no-extras > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
python bench.py --workload roundtrip512 --no-extras > gpurun_out/s6_bench512.json 2>>gpurun_out/s6_bench.err
python bench.py --no-extras --no-graphs > gpurun_out/s6_bench_nographs.json 2>>gpurun_out/s6_bench.err
python - <<EOF
import json
for f in ["s6_bench.json","s6_bench512.json","s6_bench_nographs.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["gpu_launches"])
    except Exception as e: print(f, e)
EOF
tail -5 gpurun_out/s6_bench.err
