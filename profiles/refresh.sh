#!/usr/bin/env bash
# Everything the round's committed evidence comes from, in ONE gpurun call (B200 x1):
#   gpurun --timeout 1500 -- 'bash profiles/refresh.sh r2'
# then, here:  bash profiles/collect.sh r2    (summaries of the ncu outputs + copies of the bench lines under profiles/)
# Numbers printed by the runs under ncu are never bench values.
set -uo pipefail
tag="${1:-r2}"
out=gpurun_out
mkdir -p "$out"
python -m pytest tests -m gpu -x -q > "$out/${tag}_tests.log" 2>&1; tail -2 "$out/${tag}_tests.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1; tail -1 "$out/${tag}_smoke.log"
python bench.py --steps 200 --warmup 20 > "$out/${tag}_bench_c2.json" 2> "$out/${tag}_bench_c2.err"; cut -c1-220 "$out/${tag}_bench_c2.json"
for wl in x_b512_256x256_r0.1-0.8-0.1 x_b2048_256x256_r0.1-0.8-0.1; do
    python bench.py --workload "$wl" --steps 50 --warmup 5 --no-configs --no-cpu-baseline > "$out/${tag}_bench_${wl}.json" 2> "$out/${tag}_bench_${wl}.err"; cut -c1-220 "$out/${tag}_bench_${wl}.json"
done
python bench.py --impl reference --steps 3 --warmup 1 > "$out/${tag}_bench_reference_arm.json" 2> "$out/${tag}_bench_reference_arm.err"; cut -c1-220 "$out/${tag}_bench_reference_arm.json"
# launch list of the bench command itself (per-launch times under ncu are cold-cache and serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$out/${tag}_launches.csv" python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > "$out/${tag}_ncu_launches.log" 2>&1
# one full capture of every step kernel: the 64-image batch (four launches per step) ...
ncu --set full --clock-control none --import-source on -k regex:'vq_warp_kernel|pack_kernel|unpack_decode_kernel|unpack_assemble_kernel' -s 8 -c 4 -f -o "$out/${tag}_prof" python profiles/prof_step.py > "$out/${tag}_ncu_prof.log" 2>&1
tail -1 "$out/${tag}_ncu_prof.log"
# ... and a batch that fills the machine (512 images: one-CTA-per-image packer, fused decoder)
ncu --set full --clock-control none --import-source on -k regex:'vq_warp_kernel|pack_image_kernel|unpack_small_kernel' -s 6 -c 3 -f -o "$out/${tag}_prof_b512" python profiles/prof_step.py x_b512_256x256_r0.1-0.8-0.1 4 > "$out/${tag}_ncu_prof_b512.log" 2>&1
tail -1 "$out/${tag}_ncu_prof_b512.log"
# the image-in head (entropy + routing, fine mask + mix)
ncu --set full --clock-control none --import-source on -k regex:'entropy_kernel|route_mix_kernel' -s 2 -c 2 -f -o "$out/${tag}_prof_head" python profiles/bench_image_in.py > "$out/${tag}_ncu_prof_head.log" 2>&1
tail -1 "$out/${tag}_ncu_prof_head.log"
CGIC_B200_LIB=build/variants/lib_trace.so python profiles/trace_graph.py > "$out/${tag}_trace_graph.log" 2>&1; tail -12 "$out/${tag}_trace_graph.log"
