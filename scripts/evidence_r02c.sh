set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02c_gpu_tests.txt
python bench.py > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:coset_pair_tma -s 4 -c 3 -o gpurun_out/r02c_prof_coset_pair_tma -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/r02c_gpu_tests.txt
