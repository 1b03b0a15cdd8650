#!/usr/bin/env python
"""Turns the two ncu outputs of a round into the committed summaries under profiles/:
   python profiles/summarize_ncu.py <launches.csv> <prof.ncu-rep> <tag>
-> profiles/<tag>_launches.txt (per-kernel mean duration + share of the step)
   profiles/<tag>_kernels.txt  (per-kernel dram bytes, duration, pipe utilisation from --set full)"""
import collections, csv, subprocess, sys, os
launches, rep, tag = sys.argv[1:4]
step = (sys.argv[4] if len(sys.argv) > 4 else 'vq_warp_kernel,pack_kernel<8>,unpack_decode_kernel,unpack_assemble_kernel').split(',')
suffix = sys.argv[5] if len(sys.argv) > 5 else ''
out_dir = os.path.dirname(os.path.abspath(__file__))
rows = list(csv.reader(open(launches)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr_i]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
acc = collections.OrderedDict()
other = 0
for r in rows[hdr_i + 1:]:
    if len(r) > vi:
        if 'cgic::' not in r[ki]:      # torch's own kernels in the captured command (L2 flush memset, set-up): not part of the step
            other += 1
            continue
        name = r[ki].split('(')[0].split('::')[-1]
        acc.setdefault(name, []).append(float(r[vi].replace(',', '')) / (1000 if r[ui] == 'ns' else 1))
tot = sum(sum(v) / len(v) for k, v in acc.items() if k in step)
with open(os.path.join(out_dir, f"{tag}_launches{suffix}.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): compare SHARES\n# source: {os.path.basename(launches)}\n")
    for k, v in acc.items():
        if k in step:
            f.write(f"{k:28s} launches={len(v):3d} mean_us={sum(v)/len(v):8.2f} share={sum(v)/len(v)/tot:.3f}\n")
    f.write("# other library kernels of the captured command (image-in head of bench.py: entropy / router / mask-mix; codebook index build): " +
            ", ".join(f"{k} x{len(v)} {sum(v)/len(v):.1f} us" for k, v in acc.items() if k not in step) + "\n")
    f.write(f"{'sum of kernel means':28s} {tot:.2f} us per step   ({other} launches of non-library kernels -- L2 flush, set-up -- left out)\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hh = rr[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__waves_per_multiprocessor', 'launch__cluster_size']
with open(os.path.join(out_dir, f"{tag}_kernels{suffix}.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on, one launch per kernel\n# source: {os.path.basename(rep)}\n")
    for r in rr[2:]:
        f.write(r[hh.index('Kernel Name')].split('(')[0].split('::')[-1] + "\n")
        for w in want:
            if w in hh:
                f.write(f"    {w:70s} {r[hh.index(w)]} {rr[1][hh.index(w)]}\n")
print(open(os.path.join(out_dir, f"{tag}_launches{suffix}.txt")).read())
