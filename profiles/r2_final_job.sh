set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1; tail -3 gpurun_out/r2f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -3 gpurun_out/r2f_smoke.log
python bench.py > gpurun_out/r2f_bench_encode256.json 2> gpurun_out/r2f_bench_encode256.err
python bench.py --workload roundtrip512 > gpurun_out/r2f_bench_roundtrip512.json 2> gpurun_out/r2f_bench_roundtrip512.err
python bench.py --workload slide > gpurun_out/r2f_bench_slide.json 2> gpurun_out/r2f_bench_slide.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2f_ncu_launches.log 2>&1
python profiles/run_stem.py > gpurun_out/r2f_stem_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stem_in -s 2 -c 1 -o gpurun_out/r2f_stem_in_u8 python profiles/run_stem.py > gpurun_out/r2f_stem_ncu.log 2>&1
python profiles/step_breakdown.py fp16 256 3 > gpurun_out/r2f_breakdown.txt 2>&1
cut -c1-300 gpurun_out/r2f_bench_encode256.json
