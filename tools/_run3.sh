tools/gpu_check.sh t:test_gpu_attention t:test_gpu_model t:test_gpu_fullsize
python tools/hiccup.py --steps 100 > gpurun_out/hiccup_on2.log 2>&1; tail -8 gpurun_out/hiccup_on2.log
timeout 600 python bench.py --model Cnn_9layers_Transformer_FrameAvg --batch 128 --steps 20 --no-cpu-baseline > gpurun_out/bench_transformer_r2a.log 2>&1; python tools/show_bench.py gpurun_out/bench_transformer_r2a.log | head -14
