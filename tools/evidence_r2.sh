# round-2 evidence: every file lands in gpurun_out/ and is copied (summarised) into profiles/ by hand afterwards
python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_gpu_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench.err
python tools/bench_all.py > gpurun_out/r2_all_kernels_events.txt 2>&1
python tools/train_profile.py ours 32 > gpurun_out/r2_train_step_profile.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-model --no-cpu-baseline > gpurun_out/r2_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-model --no-cpu-baseline > gpurun_out/r2_ncu1.log 2>&1
python tools/prof_small.py > gpurun_out/r2_plain4.log 2>&1 && ncu --set full --clock-control none -k regex:"dwconv|merge_norm|dt_proj|plane_transpose" -o gpurun_out/r2_small python tools/prof_small.py > gpurun_out/r2_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
