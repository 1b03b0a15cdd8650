#!/usr/bin/env bash
# Everything the round's committed evidence comes from, in ONE gpurun call (B200 x1):
#   gpurun --timeout 900 -- 'bash profiles/refresh.sh r1'
# then, here:  python profiles/summarize_ncu.py gpurun_out/<tag>_launches.csv gpurun_out/<tag>_prof.ncu-rep <tag>
#              and copy gpurun_out/<tag>_bench_*.json into profiles/<tag>/.
# Numbers printed by the runs under ncu are never bench values.
set -uo pipefail
tag="${1:-r1}"
out=gpurun_out
mkdir -p "$out"
python -m pytest tests -m gpu -x -q > "$out/${tag}_tests.log" 2>&1; tail -2 "$out/${tag}_tests.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1; tail -1 "$out/${tag}_smoke.log"
python bench.py --steps 200 --warmup 20 > "$out/${tag}_bench_c2.json" 2> "$out/${tag}_bench_c2.err"; cut -c1-220 "$out/${tag}_bench_c2.json"
for wl in c3_b24_512x768_r0.3-0.6-0.1 c3_b24_512x768_r0.1-0.8-0.1 c3_b24_512x768_r0.05-0.05-0.9; do
    python bench.py --workload "$wl" --steps 100 --warmup 10 > "$out/${tag}_bench_${wl}.json" 2> "$out/${tag}_bench_${wl}.err"; cut -c1-220 "$out/${tag}_bench_${wl}.json"
done
python bench.py --impl reference --steps 3 --warmup 1 > "$out/${tag}_bench_reference_arm.json" 2> "$out/${tag}_bench_reference_arm.err"; cut -c1-220 "$out/${tag}_bench_reference_arm.json"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$out/${tag}_launches.csv" python bench.py --steps 20 --warmup 3 > "$out/${tag}_ncu_launches.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:'vq_warp_kernel|pack_kernel|unpack_decode_kernel|unpack_assemble_kernel' -s 8 -c 4 -f -o "$out/${tag}_prof" python profiles/prof_step.py > "$out/${tag}_ncu_prof.log" 2>&1
tail -2 "$out/${tag}_ncu_prof.log"
