# run 3: the kernels changed after run 2 -- chained decoder (hand-over tables, ticket sweep), fused decoder (lengths overlaid on
# fn rows, 3 CTAs per SM), search kernel (cp.async record staging, branch-free classification), session narrow-wire kernels
mkdir -p gpurun_out
SEL="test_large_grid_all_modes_vs_oracle or test_large_grid_corrupt_and_long_codes or test_session_host_roundtrip or test_big_golden"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r2_san3_memcheck.log 2>&1
echo "== memcheck: exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_san3_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r2_san3_racecheck.log 2>&1
echo "== racecheck: exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_san3_racecheck.log | tail -3; grep -E "=========     at |Potential|Race reported" gpurun_out/r2_san3_racecheck.log | sort | uniq -c | sort -rn | head -12
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_small_grid_fused_decoder and kat5 and (256-256-0.1 or 64-48 or 256-256-0.0-0.0)" > gpurun_out/r2_san3_racecheck_fused.log 2>&1
echo "== racecheck fused decoder: exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_san3_racecheck_fused.log | tail -3; grep -E "=========     at |Potential|Race reported" gpurun_out/r2_san3_racecheck_fused.log | sort | uniq -c | sort -rn | head
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_large_grid_all_modes_vs_oracle or (test_small_grid_fused_decoder and kat5 and 256-256-0.1)" > gpurun_out/r2_san3_synccheck.log 2>&1
echo "== synccheck: exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_san3_synccheck.log | tail -3
