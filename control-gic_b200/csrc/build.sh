#!/usr/bin/env bash
# Builds libcgic_b200.so in-tree (next to the Python package) for sm_100a only.
# nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libcgic_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS=(table.cu vq_assign.cu codebook.cu pack.cu unpack.cu router.cu entropy.cu session.cu spatial_norm.cu)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --ftz=false --prec-div=true --prec-sqrt=true
       --fmad=false -Xcompiler -fPIC,-O2,-fvisibility=hidden -shared -cudart shared)
if [[ "${CGIC_PTXAS_V:-0}" == "1" ]]; then FLAGS+=(-Xptxas -v); fi
if [[ -n "${CGIC_EXTRA_FLAGS:-}" ]]; then FLAGS+=(${CGIC_EXTRA_FLAGS}); fi
OUT="${CGIC_OUT:-${OUT}}"
cd "${HERE}"
"${NVCC}" "${FLAGS[@]}" -o "${OUT}" "${SRCS[@]}"
echo "built ${OUT}"
