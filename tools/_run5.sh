tools/gpu_check.sh t:test_gpu_logmel t:test_gpu_conv_c1 t:test_gpu_fullsize t:test_gpu_model kmem
timeout 600 python bench.py --workload logmel --steps 20 --no-cpu-baseline > gpurun_out/bench_logmel_r2b.log 2>&1; tail -c 1500 gpurun_out/bench_logmel_r2b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel -s 2 -c 1 -f -o gpurun_out/logmel_prof2 python tools/prof_logmel.py > gpurun_out/ncu_logmel2.log 2>&1; tail -2 gpurun_out/ncu_logmel2.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2c.log 2>&1; python tools/show_bench.py gpurun_out/bench_r2c.log 2>/dev/null | head -16
