#!/bin/bash
# Run under gpurun: per-file GPU tests (isolated processes so one sticky CUDA error does not hide
# the rest), smoke, kernel micro-bench, bench, ncu launch list.  Logs land in gpurun_out/.
#   tools/gpu_check.sh [what...]   what in: tests smoke kbench bench ncu_list t:<test file stem> py:<script>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WHAT="${@:-tests smoke kbench bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
nproc > gpurun_out/nproc.txt
run_test() {
  n=$(basename $1 .py)
  timeout 900 python -m pytest $1 -m gpu -q --timeout 300 > gpurun_out/$n.log 2>&1
  echo "$n exit $?" | tee -a gpurun_out/summary.txt
  tail -6 gpurun_out/$n.log
}
for w in $WHAT; do
  case $w in
    tests)
      for f in tests/test_gpu_*.py; do run_test $f; done;;
    t:*)
      run_test tests/${w#t:}.py;;
    py:*)
      s=${w#py:}; n=$(basename $s .py)
      timeout 900 python $s > gpurun_out/$n.log 2>&1
      echo "$n exit $?" | tee -a gpurun_out/summary.txt; tail -60 gpurun_out/$n.log;;
    smoke)
      timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
      echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/smoke.log;;
    kbench)
      timeout 600 python tools/bench_kernels.py --batch 64 > gpurun_out/kbench.log 2>&1
      echo "kbench exit $?" | tee -a gpurun_out/summary.txt; tail -40 gpurun_out/kbench.log;;
    umma_shift)
      timeout 120 tools/experiments/umma_shift_test > gpurun_out/umma_shift.log 2>&1
      echo "umma_shift exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/umma_shift.log;;
    ncu_gru)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_.*persistent -s 4 -c 2 \
          -f -o gpurun_out/gru_prof python tools/prof_gru.py > gpurun_out/ncu_gru.log 2>&1
      echo "ncu_gru exit $?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/ncu_gru.log;;
    kbench256)
      timeout 600 python tools/bench_kernels.py --batch 256 --what conv > gpurun_out/kbench256.log 2>&1
      echo "kbench256 exit $?" | tee -a gpurun_out/summary.txt; tail -40 gpurun_out/kbench256.log;;
    kmem)
      timeout 600 python tools/bench_kernels.py --batch 256 --what bn,c1,lin > gpurun_out/kmem.log 2>&1
      echo "kmem exit $?" | tee -a gpurun_out/summary.txt; tail -40 gpurun_out/kmem.log;;
    ncu_conv)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 \
          -f -o gpurun_out/conv_prof python tools/prof_conv.py 256 > gpurun_out/ncu_conv.log 2>&1
      echo "ncu_conv exit $?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/ncu_conv.log;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench.log 2>&1
      echo "bench exit $?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/bench.log;;
    ncu_list)
      timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
          --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
          > gpurun_out/ncu_list.log 2>&1
      echo "ncu_list exit $?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/ncu_list.log;;
    *) echo "unknown target $w";;
  esac
done
exit 0
