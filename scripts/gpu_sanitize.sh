#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the hot kernels: fused decode, the TMA / cp.async layer kernels
# of one train forward + backward, the fused step and the many-block latent head.  Logs go to gpurun_out/.
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_decode_fused_A tests/test_gpu_parity.py::test_train_forward_backward_A tests/test_gpu_parity.py::test_fused_latent_head_many_blocks_is_deterministic tests/test_gpu_step.py"
for tool in memcheck racecheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest $T -x -q -p no:cacheprovider ) > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_$tool.log | tail -5
done
