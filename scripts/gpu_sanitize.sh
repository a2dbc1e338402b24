#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the hot kernels: the warp-specialised fused decode head (both
# configurations: mbarrier full/empty hand-off, cp.async weight ring, named barriers), the polyphase stem kernel,
# the TMA / cp.async layer kernels of one train forward + backward, the fused step and the many-block latent head.
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_decode_fused_A tests/test_gpu_parity.py::test_decode_layerwise_B tests/test_gpu_parity.py::test_decode_many_blocks_deterministic_and_batch_invariant tests/test_gpu_parity.py::test_train_forward_backward_A tests/test_gpu_parity.py::test_fused_latent_head_many_blocks_is_deterministic tests/test_gpu_step.py"
for tool in memcheck racecheck; do
  ( time timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest $T -x -q -p no:cacheprovider ) > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_$tool.log | tail -5
done
