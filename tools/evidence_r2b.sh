python -m pytest tests/test_ss2d_gpu.py tests/test_model_gpu.py -m gpu -q > gpurun_out/r2j_tests.log 2>&1; tail -2 gpurun_out/r2j_tests.log
python tools/bench_all.py > gpurun_out/r2j_all.txt 2>&1; grep -i "dwconv" gpurun_out/r2j_all.txt | grep "train\|micro"
python tools/fused_micro.py > gpurun_out/r2_plain3.log 2>&1 && ncu --set full --clock-control none -k regex:"plane_transpose|sl_" -s 4 -c 8 -o gpurun_out/r2_fused_micro python tools/fused_micro.py > gpurun_out/r2_ncu3.log 2>&1
tail -2 gpurun_out/r2_ncu3.log
