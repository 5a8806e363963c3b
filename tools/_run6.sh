tools/gpu_check.sh t:test_gpu_logmel t:test_gpu_gru t:test_gpu_model t:test_gpu_fullsize
timeout 600 python bench.py --workload logmel --steps 30 --no-cpu-baseline > gpurun_out/bench_logmel_r2c.log 2>&1; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_logmel_r2c.log') if l.startswith('{')][-1]);print('main  logmel', d['value'], d['ms_per_step'], d['roofline']['frac'])"
export SED_B200_LIB=$PWD/sound_event_detection_dcase2017_task4_b200/libsedb200_lm4.so
timeout 600 python -m pytest tests/test_gpu_logmel.py -m gpu -q 2>&1 | tail -2
timeout 600 python bench.py --workload logmel --steps 30 --no-cpu-baseline > gpurun_out/bench_logmel_r2c_lm4.log 2>&1; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_logmel_r2c_lm4.log') if l.startswith('{')][-1]);print('lm4   logmel', d['value'], d['ms_per_step'], d['roofline']['frac'])"
unset SED_B200_LIB
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2d.log 2>&1; python tools/show_bench.py gpurun_out/bench_r2d.log 2>/dev/null | head -24
