set -x
python bench.py > gpurun_out/r2b_bench_encode256.json 2> gpurun_out/r2b_bench_encode256.err
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2b_ncu_launches.log 2>&1
python profiles/run_resident.py 54 3 > gpurun_out/r2b_res_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trunk_resident -s 1 -c 1 -o gpurun_out/r2b_trunk_resident python profiles/run_resident.py 54 3 > gpurun_out/r2b_res_ncu.log 2>&1
python profiles/run_lowc.py 16 128 mma 4 > gpurun_out/r2b_c16_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:same_block_mma -s 2 -c 1 -o gpurun_out/r2b_mma_c16 python profiles/run_lowc.py 16 128 mma 4 > gpurun_out/r2b_c16_ncu.log 2>&1
python profiles/run_lowc.py 64 32 split 4 > gpurun_out/r2b_split_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:same_block_split -s 2 -c 1 -o gpurun_out/r2b_split_c64 python profiles/run_lowc.py 64 32 split 4 > gpurun_out/r2b_split_ncu.log 2>&1
cat gpurun_out/r2b_res_plain.log gpurun_out/r2b_c16_plain.log gpurun_out/r2b_split_plain.log | tail -12
