mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool: stem window-view conv / wgrad, uint8 staging, fused SPPF pooling (train + inference form), one-launch BN fold (eval forward), sub-module plans" 
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_conv_gpu.py tests/test_wgrad_gpu.py tests/test_elementwise_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "stem_view or 3x1 or prep_input or sppf or upsample_add_prep or forward_eval or submodule" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|error" | tail -4
  echo "exit=$?"
done > gpurun_out/sanitizer_r2z.txt 2>&1
cat gpurun_out/sanitizer_r2z.txt
