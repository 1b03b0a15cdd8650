#!/usr/bin/env bash
# After `gpurun -- bash profiles/refresh.sh <tag>`: turns gpurun_out/<tag>_* into the committed summaries under profiles/.
set -uo pipefail
tag="${1:-r2}"
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p "profiles/${tag}"
python profiles/summarize_ncu.py "gpurun_out/${tag}_launches.csv" "gpurun_out/${tag}_prof.ncu-rep" "${tag}" 'vq_warp_kernel,pack_kernel<8>,unpack_decode_kernel,unpack_assemble_kernel'
python profiles/summarize_ncu.py "gpurun_out/${tag}_launches.csv" "gpurun_out/${tag}_prof_b512.ncu-rep" "${tag}" 'vq_warp_kernel,pack_kernel<8>,unpack_decode_kernel,unpack_assemble_kernel' _b512 > /dev/null
python profiles/summarize_ncu.py "gpurun_out/${tag}_launches.csv" "gpurun_out/${tag}_prof_head.ncu-rep" "${tag}" 'vq_warp_kernel,pack_kernel<8>,unpack_decode_kernel,unpack_assemble_kernel' _head > /dev/null
rm -f "profiles/${tag}_launches_b512.txt" "profiles/${tag}_launches_head.txt"
cp "gpurun_out/${tag}_launches.csv" "profiles/${tag}/launches.csv"
for f in gpurun_out/${tag}_bench_*.json; do cp "$f" "profiles/${tag}/$(basename "$f" | sed "s/^${tag}_//")"; done
cp "gpurun_out/${tag}_trace_graph.log" "profiles/${tag}/trace_graph.txt"
tail -3 "gpurun_out/${tag}_tests.log" > "profiles/${tag}/gpu_tests.txt"; tail -1 "gpurun_out/${tag}_smoke.log" >> "profiles/${tag}/gpu_tests.txt"
ls profiles/${tag}
