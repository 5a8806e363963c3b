python tools/bench_kernels.py --batch 256 --what bn 2>&1 | grep -E "bn_bwd" | cut -c1-200
echo "== u8"; SED_B200_LIB=$PWD/sound_event_detection_dcase2017_task4_b200/libsedb200_u8.so python tools/bench_kernels.py --batch 256 --what bn 2>&1 | grep -E "bn_bwd" | cut -c1-200
tools/gpu_check.sh tests smoke
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2e.log 2>&1; python tools/show_bench.py gpurun_out/bench_r2e.log 2>/dev/null | head -30
cat gpurun_out/summary.txt
