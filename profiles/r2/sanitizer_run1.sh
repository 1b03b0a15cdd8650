# compute-sanitizer over the pack / unpack / router / fused kernels (a subset of the GPU tests: the tools slow kernels down ~50x)
SEL="test_e2e_golden or test_router_golden or test_router_ties_and_large or test_small_grid_fused_decoder_truncated or test_f1_two_launch_tail_golden or test_huffman_stream_kats or test_binary_stream_kats"
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r2_san_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_san_$tool.log | tail -3
done
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_small_grid_fused_decoder and (kat5 or short) and (256-256-0.1 or 64-48 or 16-16)" > gpurun_out/r2_san_memcheck_fused.log 2>&1
echo "== memcheck fused decoder (clusters): exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_san_memcheck_fused.log | tail -2
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_encode_fused_matches_two_launch_path and (256-256-0.1 or 64-48)" > gpurun_out/r2_san_racecheck_encode.log 2>&1
echo "== racecheck fused encoder: exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_san_racecheck_encode.log | tail -2
