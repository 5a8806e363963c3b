tools/gpu_check.sh t:test_gpu_bnpool t:test_gpu_reference_loop t:test_gpu_prep
tools/ab_multi.sh "engine.OVERLAP_LAYERS=();(2,3,4,5,6,7);(1,2,3,4,5,6,7);(3,5,7);(2,4,6)"
SED_B200_LIB=$PWD/sound_event_detection_dcase2017_task4_b200/libsedb200_ldg.so timeout 600 python tools/ab_step.py "engine.OVERLAP_LAYERS=();(2,3,4,5,6,7);(1,2,3,4,5,6,7)" --steps 20 --rounds 2 > gpurun_out/ab_ldg.log 2>&1; grep round gpurun_out/ab_ldg.log
tail -40 gpurun_out/test_gpu_reference_loop.log
