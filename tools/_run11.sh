tools/gpu_check.sh t:test_gpu_conv t:test_gpu_bnpool t:test_gpu_fullsize t:test_gpu_model
timeout 600 python tools/bench_kernels.py --batch 256 --what conv --pair > gpurun_out/kbench256_r2.log 2>&1; grep -E "wgrad" gpurun_out/kbench256_r2.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2f.log 2>&1; python tools/show_bench.py gpurun_out/bench_r2f.log 2>/dev/null | head -26
