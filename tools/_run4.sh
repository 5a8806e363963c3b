tools/gpu_check.sh t:test_gpu_conv_c1 t:test_gpu_attention t:test_gpu_model t:test_gpu_fullsize kmem
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2b.log 2>&1; python tools/show_bench.py gpurun_out/bench_r2b.log 2>/dev/null | head -16
timeout 600 python bench.py --model Cnn_9layers_Transformer_FrameAvg --batch 128 --steps 20 --no-cpu-baseline > gpurun_out/bench_transformer_r2b.log 2>&1; python tools/show_bench.py gpurun_out/bench_transformer_r2b.log 2>/dev/null | head -12
