tools/gpu_check.sh t:test_gpu_reference_loop
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel -s 2 -c 1 -f -o gpurun_out/logmel_prof python tools/prof_logmel.py > gpurun_out/ncu_logmel.log 2>&1; tail -2 gpurun_out/ncu_logmel.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.log 2>&1; tail -c 1500 gpurun_out/bench_r2a.log
timeout 600 python bench.py --workload logmel --steps 20 > gpurun_out/bench_logmel_r2a.log 2>&1; tail -c 2500 gpurun_out/bench_logmel_r2a.log
timeout 600 python bench.py --workload eval --steps 5 --batch 64 > gpurun_out/bench_eval_r2a.log 2>&1; tail -c 1200 gpurun_out/bench_eval_r2a.log
