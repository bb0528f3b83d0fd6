#!/bin/bash
# Runs the op-level GPU parity tests in separate processes (a faulting kernel poisons its CUDA
# context) with per-group timeouts; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in "gemm_plain" "gemm_epilogues or gemm_scatter or gemm_broadcast or gemm_blockln32" "layernorm or softmax or poswise or opm_prep or pair2att or instnorm or launch_counter" "favor"; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-40)
  timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$grp" --timeout 120 --tb=short > "gpurun_out/ops_$name.log" 2>&1
  echo "== $grp: exit $?"; tail -n 30 "gpurun_out/ops_$name.log"
done
timeout 900 python -m pytest tests/test_gpu_trunk.py -q -m gpu --timeout 600 --tb=short -s > gpurun_out/trunk.log 2>&1
echo "== trunk: exit $?"; grep -E "rel-l2|chain|forced|passed|failed|Error|error" gpurun_out/trunk.log | tail -n 40
