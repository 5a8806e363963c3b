python tools/bench_kernels.py --batch 256 --what bn 2>&1 | grep -E "pool\": 2" | cut -c1-200
for v in u4 u1; do echo "== $v"; SED_B200_LIB=$PWD/sound_event_detection_dcase2017_task4_b200/libsedb200_$v.so python tools/bench_kernels.py --batch 256 --what bn 2>&1 | grep -E "pool\": 2" | cut -c1-200; done
