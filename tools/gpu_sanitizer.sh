#!/bin/bash
# compute-sanitizer over the hand-written mbarrier / TMEM / TMA pipelines: conv (single-CTA and CTA-pair, bf16 and fp16),
# wgrad, fused Fcomb forward / backward.  ONE tool per gpurun call (TOOL=memcheck|racecheck|synccheck|initcheck).
mkdir -p gpurun_out
TOOL=${TOOL:-memcheck}
SEL='test_conv3x3_matches_fp32_reference or cta_pair or test_conv3x3_pool_only or test_first_conv or test_avgpool_and_upsample or test_conv3x3_fp16_range'
timeout 1700 compute-sanitizer --tool $TOOL --target-processes all --error-exitcode 7 --log-file gpurun_out/sanitizer_${TOOL}_conv.log \
  python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 1500 -p no:cacheprovider -x -k "$SEL" > gpurun_out/sanitizer_${TOOL}_conv.pytest.log 2>&1
echo "conv rc=$?"; tail -3 gpurun_out/sanitizer_${TOOL}_conv.pytest.log; tail -4 gpurun_out/sanitizer_${TOOL}_conv.log
timeout 1700 compute-sanitizer --tool $TOOL --target-processes all --error-exitcode 7 --log-file gpurun_out/sanitizer_${TOOL}_fcomb_wgrad.log \
  python -m pytest tests/test_gpu_punet.py tests/test_gpu_train.py -m gpu -q --timeout 1500 -p no:cacheprovider -x \
  -k "test_fcomb_kernel_parity or test_fcomb_sample_counts or test_conv3x3_backward or test_fcomb_backward or test_first_conv_backward" > gpurun_out/sanitizer_${TOOL}_fcomb_wgrad.pytest.log 2>&1
echo "fcomb/wgrad rc=$?"; tail -3 gpurun_out/sanitizer_${TOOL}_fcomb_wgrad.pytest.log; tail -4 gpurun_out/sanitizer_${TOOL}_fcomb_wgrad.log
