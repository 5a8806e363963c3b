#!/bin/bash
# A/B runs of the training step under the module-level schedule switches (run under gpurun).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "$@"; do
  n=$(echo "$spec" | tr '.=,' '___')
  timeout 600 python tools/ab_step.py "$spec" --steps 20 --rounds 3 > gpurun_out/ab_$n.log 2>&1
  echo "ab $spec exit $?"; grep round gpurun_out/ab_$n.log
done
