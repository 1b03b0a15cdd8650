#!/usr/bin/env python
"""bench.py -- Mpixels/s of the VQ + entropy-coding hot path (encode + decode) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic images:
    encode = VectorQuantize2.forward (a1) + index selection + 5-stream Huffman/binary pack (a7,a9,a11,a12)   cgic_encode
    decode = 5-stream unpack + mask/index re-assembly + codebook gather (a10,a11,a13,a14)                    cgic_unpack
Headline workload at every N: 64 images of 256x256 per GPU (BASELINE.json configs[1]; configs[3] is the same thing at
N = 8), ratio (0.1, 0.8, 0.1), K = 1024 codebook; images are sharded contiguously over ranks with no data-path
collective, one all-reduce of {bytes, pixels, sqerr} at the end ("scaling": "weak").

    value     device-resident: inputs already in HBM, each step timed by a CUDA-event pair on the launching stream, L2
              flushed (256 MiB memset) between steps outside the pairs.
    e2e       the host-buffer C-ABI round trip (= CGIC.compress, model.py:206-401: encode + pack + unpack + re-assembly):
              pinned host inputs -> H2D -> kernels -> D2H of every result (streams, sizes, decoded indices / masks /
              latents), wall clock per step, two round trips in flight (cgic_session_roundtrip_arena_submit / _wait on the
              session's two arena sets); `blocking_call_ms` is the single blocking call cgic_session_roundtrip_arena.  `e2e_decoded_on_device` is the same call leaving the decoded
              tensors in HBM for the decoder CNN (what model.py:391-399 does): only streams, sizes, status come back.
              `e2e_narrow_wire`: every result comes back, but masks travel as bytes and decoded indices as int16.
    configs   the other BASELINE configs through the same step: configs[2] (24 x 512x768 per GPU at its three ratios),
              configs[3] as STRONG scaling (512 images of 256x256 split over the ranks), configs[4] (the six tiles of one
              2032x1344 image, tiles dealt round-robin over the ranks).
    roofline  dominant kernel, its duration measured live with the library's per-launch CUDA events (cgic_prof_*),
              against the algorithmic bytes of DESIGN.md and the measured HBM peak of MEASURED_PEAKS.json.
    cpu_baseline  the reference's own classes (baseline/_ref, kind "reference"; the port of oracle/refport.py when that
              copy is absent) on a bounded sample of the same images, on this box's host cores (rank 0, N = 1 only), with
              a parity gate: indices and all five streams byte-identical to the GPU's.
    rank_identity (N > 1) rank 0 re-runs every other rank's image range in its own process and compares stream sizes and a
              sha256 of every image's five streams with what the ranks produced.
--impl reference times the CPU arm alone and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mpixels/s encode+decode (VQ+entropy)"
UNIT = "Mpixels/s"
WORKLOADS = {
    # name: (images per GPU, H, W, coarse ratio, medium ratio)
    "c2_b64_256x256_r0.1-0.8-0.1": (64, 256, 256, 0.1, 0.8),
    "c3_b24_512x768_r0.3-0.6-0.1": (24, 512, 768, 0.3, 0.6),
    "c3_b24_512x768_r0.1-0.8-0.1": (24, 512, 768, 0.1, 0.8),
    "c3_b24_512x768_r0.05-0.05-0.9": (24, 512, 768, 0.05, 0.05),
    # not BASELINE configs: the c2 workload at larger batches (how the kernels behave once the grid fills the machine)
    "x_b512_256x256_r0.1-0.8-0.1": (512, 256, 256, 0.1, 0.8),
    "x_b2048_256x256_r0.1-0.8-0.1": (2048, 256, 256, 0.1, 0.8),
    "x_b2_768x768_r0.1-0.8-0.1": (2, 768, 768, 0.1, 0.8),          # the largest tile group of config 5 on its own
}
DEFAULT_WORKLOAD = "c2_b64_256x256_r0.1-0.8-0.1"
C3 = ("c3_b24_512x768_r0.3-0.6-0.1", "c3_b24_512x768_r0.1-0.8-0.1", "c3_b24_512x768_r0.05-0.05-0.9")
C4_IMAGES = 512                      # BASELINE configs[3]: batch = 512 of 256x256 split over the ranks (strong scaling)
C5_IMAGE = (2032, 1344)              # BASELINE configs[4]: DIV2K 2040x1356 after the reference's crop to multiples of 16
SEED = 1000

_SAMPLER = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
out = open(sys.argv[2], "w")
while True:
    try:
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    out.write("%.6f %d %d %d %.1f\n" % (time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, r,
                                        nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
    out.flush()
    time.sleep(0.002)
"""
_REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
            0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}


class ClockSampler:
    """nvml samples of SM clock / throttle reasons in a side process (no GIL contention with the launch loop)."""

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="cgic_clocks_")
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER, str(gpu_index), self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [l.split() for l in open(self.path) if l.strip()]
            os.unlink(self.path)
        except OSError:
            return None
        rows = [(float(a), int(b), int(c), int(d), float(e)) for a, b, c, d, e in (r for r in rows if len(r) == 5)]
        if not rows:
            return None
        inside = [r for r in rows if t0 <= r[0] <= t1] or rows
        bits = 0
        for r in inside:
            bits |= r[3]
        reasons = [n for b, n in _REASONS.items() if bits & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": inside[0][2], "reasons": reasons,
                "samples": len(inside), "power_w_max": max(r[4] for r in inside)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline workload only (skip the configs[2..4] lines)")
    ap.add_argument("--e2e-parts", type=int, default=0, help="image ranges the pinned-arena round trip is cut in (0 = 4)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    return ap.parse_args()


def config_of(args, world):
    B, H, W, c, m = WORKLOADS[args.workload]
    return {"workload": args.workload, "images_per_gpu": B, "global_images": B * world, "height": H, "width": W,
            "ratio": [c, m, round(1 - c - m, 6)], "codebook": 1024, "parallelism": f"image-sharded dp{world}",
            "latents": "codebook[randint] + 1e-4 * N(0,1), mixed per granularity mask (SURVEY.md 8d)",
            "l2": "flushed between timed steps (256 MiB memset outside the event pairs)"}


def stream_digests(packed, sizes, offs):
    """sha256 over the five valid streams of every image (host side) -> list of 32-byte digests."""
    blob, sz = packed.cpu().numpy(), sizes.cpu().numpy()
    out = []
    for b in range(blob.shape[0]):
        hsh = hashlib.sha256()
        for s in range(5):
            hsh.update(blob[b, offs[s]: offs[s] + sz[b, s]].tobytes())
        out.append(hsh.digest())
    return out


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own classes (or the reference-shaped port) on host cores
# ---------------------------------------------------------------------------------------------
def cpu_inputs(B, H, W, c, m, seed, first=0):
    """Same seeded inputs as the GPU arm, built on the CPU with the oracle's router / mask-mix."""
    import numpy as np
    import torch

    import workload
    from oracle import oracle as orc
    cbk, counts = workload.codebook_and_counts()
    e16, e8 = workload.entropy_maps(B, H, W, seed, first)
    hc, hm, hf = workload.heads(B, H, W, cbk, seed, first)
    zs, masks = [], []
    for b in range(B):
        mc, mm, mf, mode = orc.router(e16[b:b + 1].numpy(), e8[b:b + 1].numpy(), c, m)
        z = orc.mask_mix(hc[b:b + 1].numpy(), hm[b:b + 1].numpy(), hf[b:b + 1].numpy(), mc, mm, mf)
        zs.append(torch.from_numpy(z))
        masks.append(tuple(torch.from_numpy(np.ascontiguousarray(t)) for t in (mc, mm, mf)))   # [1,1,.,.] like grain_mask
    return cbk, counts, zs, masks


def cpu_run(args, n_images_cap, seconds, steps=None, warmup=0, verify=False):
    """Times the CPU arm's round trip (VQ -> select -> 5 files -> read back -> re-assemble -> gather) per image.
    steps=None: repeat passes over the sample until `seconds` have elapsed.  verify: one more, untimed pass that
    returns every sample image's indices and the sha256 of its five files (the parity gate's CPU side)."""
    import torch

    import workload
    from oracle import refarm
    B, H, W, c, m = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = min(B, n_images_cap)
    cbk, counts, zs, masks = cpu_inputs(n, H, W, c, m, seed=SEED)
    arm = refarm.Arm(cbk, counts.tolist(), workload.lexicographic_order())
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    sizes_all, check = [], []
    with tempfile.TemporaryDirectory(dir=shm) as tmp:
        def one_pass():
            return [arm.roundtrip(zs[b], masks[b], tmp)[4] for b in range(n)]
        for _ in range(warmup):
            one_pass()
        t0 = time.perf_counter()
        passes = 0
        while True:
            sizes_all = one_pass()
            passes += 1
            if steps is not None and passes >= steps:
                break
            if steps is None and time.perf_counter() - t0 >= seconds:
                break
        dt = time.perf_counter() - t0
        if verify:
            for b in range(n):
                ind, bpp, ind_dec, quant, sizes = arm.roundtrip(zs[b], masks[b], tmp)
                hsh = hashlib.sha256()
                for data in arm.files(tmp):
                    hsh.update(data)
                check.append((ind.flatten().clone(), hsh.digest(), ind_dec.flatten().clone()))
    mpix = passes * n * H * W / 1e6
    what = ("the reference's own VectorQuantize2 / HuffmanCoding / BinaryCoding (baseline/_ref)" if arm.kind == "reference"
            else "reference-shaped port (oracle/refport.py)")
    return {"value": mpix / dt, "unit": UNIT, "cores": cores, "kind": arm.kind,
            "sample": f"{n} of the {B} images of one step x {passes} passes, {dt:.1f} s, {what}, torch threads = {cores}, files on {shm or 'tmp'}",
            "ms_per_image": 1e3 * dt / (passes * n)}, sizes_all, dt / passes, check


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    B, H, W, c, m = WORKLOADS[args.workload]
    # calibrate on one image, then size the per-step sample so that the whole run stays within ~2 minutes
    cal, _, t_img, _ = cpu_run(args, 1, 0, steps=1, warmup=1)
    budget = 120.0
    n = max(1, min(B, int(budget / max(1, args.steps + args.warmup) / max(t_img, 1e-4))))
    res, _, t_step, _ = cpu_run(args, n, 0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(args, world),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
class HotPath:
    """Replicated state of the path on one GPU: code table, codebook, its search index."""

    def __init__(self, dev):
        import cgic_b200 as cg
        import workload
        self.cg, self.dev = cg, dev
        self.cbk, self.counts = workload.codebook_and_counts()
        self.table = cg.ops.HuffTable(self.counts.numpy(), workload.lexicographic_order()).upload()
        self.cb = self.cbk.to(dev)
        self.prepared = cg.ops.Codebook(self.cb)               # VectorQuantize2's prepared codebook, built once per weight version

    def inputs(self, B, H, W, c, m, first):
        """Latents z [B,4,H/4,W/4] and router masks of global images first .. first+B-1 (seeded per global image id)."""
        import workload
        ops = self.cg.ops
        e16, e8 = workload.entropy_maps(B, H, W, SEED, first)
        mc, mm, mf, _, mode = ops.router(e16.to(self.dev), e8.to(self.dev), c, m, per_image=True)
        hc, hm, hf = (t.to(self.dev) for t in workload.heads(B, H, W, self.cbk, SEED, first))
        z = ops.mask_mix(hc, hm, hf, mc, mm, mf)
        return z, (mc, mm, mf), mode

    def step_fn(self, groups):
        """groups: list of (z, masks, mode).  One step = for every group: cgic_encode, then cgic_unpack.  Several groups (the
        tile shapes of one tiled image) are independent: each runs on a stream of its own, forked from and joined to the
        launching stream, so under graph capture they become parallel branches."""
        import torch
        ops = self.cg.ops
        side = [torch.cuda.Stream() for _ in groups[1:]]

        def one(z, masks, mode):
            mc, mm, mf = masks
            h, w = z.shape[-2:]
            idx, zq, sq, packed, sizes = ops.encode(z, self.prepared, mc, mm, mf, mode, self.table)
            dmc, dmm, dmf, ind, quant, status = ops.unpack(packed, sizes, mode, self.table, self.cb, h, w)
            return idx, sq, packed, sizes, ind, quant, status

        def step():
            if len(groups) <= 1:
                return [one(*g) for g in groups]
            main = torch.cuda.current_stream()
            outs = [None] * len(groups)
            for k, st in enumerate(side):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    outs[k + 1] = one(*groups[k + 1])
            outs[0] = one(*groups[0])
            for st in side:
                main.wait_stream(st)
            return outs
        return step


def launches_per_step(lib, step, flush, reps=4):
    """Kernel launches of the library inside one step, counted from its own per-launch records."""
    import ctypes
    lib.cgic_prof_enable(1)
    for _ in range(reps):
        step()
    buf = ctypes.create_string_buffer(8192)
    n = lib.cgic_prof_report(buf, 8192)
    lib.cgic_prof_enable(0)
    assert n >= 0
    return sum(int(line.split()[1]) for line in buf.value.decode().splitlines()) // reps


def capture(torch, step, no_graph):
    if no_graph:
        return step, None, None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):      # the warm-up stream: its workspaces exist, nothing but the step's kernels is captured
        g_out = step()
    return graph.replay, graph, g_out


def time_steps(torch, dist, runner, flush, steps, warmup, world, dev):
    """-> (sum of per-step device ms, MAX over ranks)."""
    for _ in range(max(warmup, 3)):
        flush.zero_()
        runner()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_()
        a.record()
        runner()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def check_roundtrip(torch, outs):
    for idx, sq, packed, sizes, ind, quant, status in outs:
        assert int(status.abs().sum()) == 0 and int(sizes.min()) >= 0
        assert torch.equal(ind.view(-1), idx), "decode(encode(idx)) != idx"


def run_config(torch, dist, hp, groups, pixels_global, args, world, dev, flush, steps):
    """One extra BASELINE config through the same step -> dict(value, ms_per_step, bytes)."""
    step = hp.step_fn(groups)
    outs = step()
    torch.cuda.synchronize()
    check_roundtrip(torch, outs)
    stream_bytes = float(sum(int(o[3].sum()) for o in outs))
    runner, graph, _ = capture(torch, step, args.no_graph or not groups)
    ms = time_steps(torch, dist, runner, flush, steps, min(args.warmup, 10), world, dev)
    t = torch.tensor([stream_bytes, float(sum(algorithmic_step_bytes(z.shape[0], z.shape[-2], z.shape[-1], float(o[3].sum()))
                                              for (z, _, _), o in zip(groups, outs)))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    tot_bytes, alg = t.tolist()
    per_step = ms / steps * 1e-3
    return {"value": pixels_global * 1e-6 / per_step, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
            "bpp": 8.0 * tot_bytes / pixels_global, "step_algorithmic_bytes": alg,
            "step_frac": alg / world / per_step / 1e9 / hbm_peak()[0]}


_PEAK = None


def hbm_peak():
    global _PEAK
    if _PEAK is None:
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            _PEAK = (float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)")
        except (OSError, KeyError, ValueError):
            _PEAK = (6650.0, "fallback 6650 GB/s (of fallback)")
    return _PEAK


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import cgic_b200 as cg
    import workload
    from cgic_b200 import dist as cdist
    from cgic_b200 import inference as cinf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    # stdout carries exactly ONE line, the JSON: everything libraries print meanwhile (NCCL's version banner ...) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, H, W, c, m = WORKLOADS[args.workload]
    h, w = H // 4, W // 4
    first = rank * B                  # every rank owns its own contiguous range of the global image list
    hp = HotPath(dev)
    table, cb, prepared, cbk = hp.table, hp.cb, hp.prepared, hp.cbk
    z, (mc, mm, mf), mode = hp.inputs(B, H, W, c, m, first)
    pixels = B * H * W
    lib = cg._lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    step = hp.step_fn([(z, (mc, mm, mf), mode)])
    # correctness gate of the run itself (round trip + status), before any timing
    outs = step()
    torch.cuda.synchronize()
    check_roundtrip(torch, outs)
    idx, sq, packed, sizes, ind, quant, status = outs[0]
    sizes_first = sizes.cpu()
    offs, _, _ = table.layout(h, w)
    digests = stream_digests(packed, sizes, offs)
    kernels_per_step = launches_per_step(lib, step, flush)
    n_c, n_m, n_f = workload.expected_counts(H, W, c, m)
    exhaustive = cg.ops.exhaustive_count("encode") + cg.ops.exhaustive_count("vq")
    exhaustive_frac = exhaustive / float(5 * B * (n_c + n_m + n_f))            # the step ran 5 times so far (gate + 4 counted)

    runner, graph, g_out = capture(torch, step, args.no_graph)
    sampler = ClockSampler(local) if rank == 0 and not os.environ.get('CGIC_BENCH_NO_SAMPLER') else None
    t_begin = time.time()
    dev_ms = time_steps(torch, dist, runner, flush, args.steps, args.warmup, world, dev)
    outs = [g_out[0]] if graph is not None else step()
    torch.cuda.synchronize()
    assert torch.equal(outs[0][3].cpu(), sizes_first) and torch.equal(outs[0][4].view(-1), outs[0][0])
    # the path's only collective: {bytes, pixels, squared error} summed over ranks -- timed, for the job-level figure (one
    # untimed call first: NCCL builds its channels for a new (dtype, op) pair on first use)
    cdist.reduce_rate_distortion(0.0, 0.0, 0.0, device=dev)
    ev_r = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev_r[0].record()
    tot_bytes, tot_pix, tot_sq, bpp = cdist.reduce_rate_distortion(float(sizes_first.sum()), float(pixels), float(sq.item()), device=dev)
    ev_r[1].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev_r[0].elapsed_time(ev_r[1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    reduce_ms = float(t.item())

    # ---- N > 1: what the ranks produced == what ONE process produces for the same global images (sizes + stream digests)
    rank_identity = None
    if world > 1:
        all_sizes = cdist.gather_sizes(sizes)
        dg = torch.tensor(np.frombuffer(b"".join(digests), np.uint8).reshape(B, 32).copy(), device=dev)
        all_dg = [torch.empty_like(dg) for _ in range(world)]
        dist.all_gather(all_dg, dg)
        if rank == 0:
            same_sizes, same_streams = True, True
            single_bytes = 0.0
            for r in range(world):
                if r == 0:
                    r_sizes, r_digests = sizes_first, digests
                else:
                    zr, mr, mode_r = hp.inputs(B, H, W, c, m, r * B)
                    o = hp.step_fn([(zr, mr, mode_r)])()[0]
                    torch.cuda.synchronize()
                    r_sizes, r_digests = o[3].cpu(), stream_digests(o[2], o[3], offs)
                single_bytes += float(r_sizes.sum())
                same_sizes &= bool(torch.equal(all_sizes[r * B:(r + 1) * B].cpu(), r_sizes))
                same_streams &= bytes(all_dg[r].cpu().numpy().tobytes()) == b"".join(r_digests)
            rank_identity = {"images": world * B, "sizes_equal": same_sizes, "stream_sha256_equal": same_streams,
                             "reduced_bpp_equal": 8.0 * single_bytes / (world * pixels) == bpp,
                             "how": "rank 0 re-ran every rank's global image range in its own process"}
            assert same_sizes and same_streams and rank_identity["reduced_bpp_equal"], rank_identity

    # ---- "image-in" figure: the same step preceded by a4 entropy maps + a5 router + a6 mask-mix on the images
    gimg = torch.Generator().manual_seed(SEED + 7 + first)
    x_img = torch.rand(B, 3, H, W, generator=gimg).to(dev)
    hc2, hm2, hf2 = (t_.to(dev) for t_ in workload.heads(B, H, W, cbk, SEED, first))

    def step_image_in():
        # two launches that read the image once (entropy maps + per-image routing; fine mask + mask-mix)
        e8_, e16_, mc_, mm_, near_, mode_ = cg.ops.entropy_route(x_img, c, m)
        mf_, _, z_ = cg.ops.route_mix(hc2, hm2, hf2, mc_, mm_, mode_)
        idx_, zq_, sq_, packed_, sizes_ = cg.ops.encode(z_, prepared, mc_, mm_, mf_, mode_, table)
        out_ = cg.ops.unpack(packed_, sizes_, mode_, table, cb, h, w)
        return idx_, out_[3], out_[5], near_
    i_idx, i_ind, i_status, i_near = step_image_in()
    torch.cuda.synchronize()
    assert int(i_status.abs().sum()) == 0 and torch.equal(i_ind.view(-1), i_idx)
    near_total = [int(v) for v in i_near.sum(0).tolist()]
    img_launches = launches_per_step(lib, step_image_in, flush, reps=2)
    img_runner, _, _ = capture(torch, step_image_in, args.no_graph)
    n_img = max(10, args.steps // 2)
    img_ms = time_steps(torch, dist, img_runner, flush, n_img, 3, world, dev) / n_img
    del hc2, hm2, hf2, x_img

    # ---- e2e: host buffers through the C-ABI session (H2D + kernels + D2H inside the timed region)
    sess = cg.ops.Session(B, h, w, mode, table, cbk)
    zh = z.cpu().pin_memory()
    mh = [t_.cpu().pin_memory() for t_ in (mc, mm, mf)]
    for _ in range(3):
        by, sz = sess.compress(zh, *mh)
        out = sess.decompress(by, sz)
    assert torch.equal(sz, sizes_first) and torch.equal(out[3].view(-1), idx.cpu()) and int(out[5].abs().sum()) == 0
    for _ in range(3):
        rt = sess.roundtrip(zh, *mh)
    assert torch.equal(rt[1], sizes_first) and torch.equal(rt[5].view(-1), idx.cpu()) and int(rt[7].abs().sum()) == 0
    # the same call on the session's pinned arenas: one copy per direction and image range, ranges pipelined, CUDA graph
    n_e2e = max(10, args.steps)
    e2e, e2e_block, e2e_parts = {}, {}, {}
    # Two round trips in flight (two arena sets, submit / wait): while one batch's results travel back the next one's
    # inputs travel in.  4 image ranges per round trip: the most robust choice over the boxes probed (profiles/e2e_probe.py;
    # host<->device copy rates on these VMs vary by up to 2x from box to box and run to run).  The blocking single call
    # (cgic_session_roundtrip_arena: submit + wait) is timed next to it.
    for key, on_device, narrow, parts in (("full", False, False, args.e2e_parts or 4), ("narrow_wire", False, True, args.e2e_parts or 4),
                                          ("decoded_on_device", True, False, args.e2e_parts or 4)):
        kw = dict(decoded_on_device=on_device, narrow=narrow)
        sets = [sess.arena(parts, slot=k) for k in (0, 1)]
        e2e_parts[key] = len(sets[0])
        for views in sets:
            for v in views:
                r = v["images"]
                v["z"].copy_(zh[r.start:r.stop])
                for name, src in zip(("m_c8", "m_m8", "m_f8") if narrow else ("m_c", "m_m", "m_f"), mh):
                    v[name].copy_(src[r.start:r.stop].to(v[name].dtype))
                v["sizes"].zero_()
                v["ind"].zero_()
                v["ind16"].zero_()

        def piped(n):
            sess.submit_arena(0, **kw)
            for i in range(1, n):
                sess.submit_arena(i & 1, **kw)       # the next batch goes in ...
                sess.wait_arena((i - 1) & 1)         # ... while the previous one comes back
            sess.wait_arena((n - 1) & 1)

        for _ in range(3):
            sess.roundtrip_arena(**kw)
        piped(6)
        for k, views in enumerate(sets):
            assert torch.equal(torch.cat([v["sizes"] for v in views]), sizes_first) and int(sum(int(v["status"].abs().sum()) for v in views)) == 0
            if narrow:
                assert torch.equal(torch.cat([v["ind16"].reshape(-1) for v in views]).long(), idx.cpu())
                assert torch.equal(torch.cat([v["mf8"].reshape(-1) for v in views]).long(), mh[2].reshape(-1).long())
            elif not on_device:
                assert torch.equal(torch.cat([v["ind"].reshape(-1) for v in views]), idx.cpu())
            else:
                assert torch.equal(sess.device_tensor("ind", slot=k).reshape(-1), idx)   # decoded tensors stayed in HBM, and are right
        for which, fn in (("blocking", lambda: [sess.roundtrip_arena(**kw) for _ in range(n_e2e)]), ("piped", lambda: piped(n_e2e))):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            fn()                                     # = CGIC.compress n_e2e times: encode + pack + unpack + re-assembly, pinned host in / out
            e2e_s = time.perf_counter() - t0
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            (e2e_block if which == "blocking" else e2e)[key] = float(t.item())
    t_end = time.time()
    n4, n8, n16 = B * h * w, B * h * w // 4, B * h * w // 16
    blob = B * sess.image_stride
    h2d = n4 * 16 + (n4 + n8 + n16) * 4
    d2h_wire = blob + B * 20 + B * 4 + 8
    d2h = d2h_wire + n4 * 8 + n4 * 16 + (n4 + n8 + n16) * 8
    sess.close()
    clocks = sampler.stop(t_begin, t_end) if sampler else None

    # ---- roofline of the dominant kernel: per-launch CUDA events inside the library, eager launches
    prof_steps = 20
    lib.cgic_prof_enable(1)
    for _ in range(prof_steps):
        flush.zero_()
        step()
    import ctypes
    buf = ctypes.create_string_buffer(8192)
    cg._lib.check(min(lib.cgic_prof_report(buf, 8192), 0), "cgic_prof_report")
    lib.cgic_prof_enable(0)
    kern = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        kern[name] = {"launches_per_step": int(cnt) / prof_steps, "us_per_launch": 1e3 * float(ms) / int(cnt)}
    step_us = sum(k["us_per_launch"] * k["launches_per_step"] for k in kern.values())
    for k in kern.values():
        k["share"] = k["us_per_launch"] * k["launches_per_step"] / step_us
    top = max(kern, key=lambda n: kern[n]["share"])
    stream_bytes = float(sizes_first.sum())
    alg = algorithmic_bytes(B, h, w, stream_bytes)
    peak, peak_source = hbm_peak()
    achieved = alg[top] / (kern[top]["us_per_launch"] * 1e-6) / 1e9
    step_s = dev_ms / args.steps * 1e-3
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_source, "traffic": _traffic(top, B), "traffic_source": NCU_TRAFFIC_SOURCE if _traffic(top, B) is not None else None,
                "algorithmic_bytes_per_launch": alg[top],
                "us_per_launch": kern[top]["us_per_launch"], "kernels": kern,
                "step_algorithmic_bytes": alg["step"], "step_achieved_gbs": alg["step"] / step_s / 1e9, "step_frac": alg["step"] / step_s / 1e9 / peak}

    value = world * pixels * args.steps / 1e6 / (dev_ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(args, world),
            "e2e": {"value": world * pixels * n_e2e / 1e6 / e2e["full"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": n_e2e, "ms_per_step": 1e3 * e2e["full"] / n_e2e, "image_ranges": e2e_parts["full"], "in_flight": 2,
                    "call": "cgic_session_roundtrip_arena_submit / _wait on the session's two pinned arena sets; every result comes back in the "
                            "reference's types (int64 indices and masks, fp32 latents)",
                    "blocking_call_ms": 1e3 * e2e_block["full"] / n_e2e},
            "e2e_narrow_wire": {"value": world * pixels * n_e2e / 1e6 / e2e["narrow_wire"], "unit": UNIT, "h2d_bytes_per_step": n4 * 16 + (n4 + n8 + n16),
                                "d2h_bytes_per_step": d2h_wire + n4 * 2 + n4 * 16 + (n4 + n8 + n16), "ms_per_step": 1e3 * e2e["narrow_wire"] / n_e2e,
                                "image_ranges": e2e_parts["narrow_wire"], "in_flight": 2, "blocking_call_ms": 1e3 * e2e_block["narrow_wire"] / n_e2e,
                                "call": "cgic_session_roundtrip_arena(flags | 8): every result still comes back, masks as one byte per cell both ways "
                                        "and decoded indices as int16 (the reference's int32 / int64 tensors exist on the device only)"},
            "e2e_decoded_on_device": {"value": world * pixels * n_e2e / 1e6 / e2e["decoded_on_device"], "unit": UNIT, "h2d_bytes_per_step": h2d,
                                      "d2h_bytes_per_step": d2h_wire, "ms_per_step": 1e3 * e2e["decoded_on_device"] / n_e2e, "image_ranges": e2e_parts["decoded_on_device"],
                                      "in_flight": 2, "blocking_call_ms": 1e3 * e2e_block["decoded_on_device"] / n_e2e,
                                      "call": "cgic_session_roundtrip_arena(flags | 4): streams, sizes, status come back; ind / quant / masks stay "
                                              "in HBM for the decoder CNN, as in model.py:391-399"},
            "gpu_launches": kernels_per_step * args.steps, "kernels_per_step": kernels_per_step, "cuda_graph": graph is not None,
            "image_in": {"value": world * pixels / 1e6 / (img_ms * 1e-3), "unit": UNIT, "ms_per_step": img_ms, "steps": n_img,
                         "kernels_per_step": img_launches,
                         "adds": "a4 entropy maps (12 B/pixel image read) + a5 router + a6 mask-mix in front of the step, as the two launches "
                                 "cgic_entropy_route + cgic_route_mix",
                         "threshold_adjacent_cells": {"coarse": near_total[0], "medium": near_total[1],
                                                      "of": [B * (H // 16) * (W // 16), B * (H // 8) * (W // 8)]}},
            "job": {"value": world * pixels * args.steps / 1e6 / ((dev_ms + reduce_ms) * 1e-3), "unit": UNIT, "reduce_ms": reduce_ms,
                    "includes": f"the {args.steps} timed steps + the final all-reduce of {{bytes, pixels, sqerr}} (the path's only collective)"},
            "vq_exhaustive_leader_frac": exhaustive_frac,
            "bpp": bpp, "stream_bytes_per_step": tot_bytes, "clocks": clocks, "roofline": roofline}
    if rank_identity is not None:
        line["rank_identity"] = rank_identity

    # ---- the other BASELINE configs through the same step
    if not args.no_configs and args.workload == DEFAULT_WORKLOAD:
        cfg = {}
        k_steps = max(10, min(args.steps, 50))
        del z, mc, mm, mf, outs, g_out, graph, runner
        for name in C3:                                                          # configs[2]: weak, 24 Kodak-shape images per GPU
            Bc, Hc, Wc, cc, mc_ = WORKLOADS[name]
            zc, masks_c, mode_c = hp.inputs(Bc, Hc, Wc, cc, mc_, rank * Bc)
            cfg[name] = dict(run_config(torch, dist, hp, [(zc, masks_c, mode_c)], world * Bc * Hc * Wc, args, world, dev, flush, k_steps),
                             scaling="weak", images_per_gpu=Bc)
            del zc, masks_c
        lo, hi = cdist.shard_range(C4_IMAGES, rank, world)                        # configs[3]: strong, 512 images split over the ranks
        zc, masks_c, mode_c = hp.inputs(hi - lo, 256, 256, 0.1, 0.8, lo)
        cfg["c4_b512_256x256_r0.1-0.8-0.1"] = dict(run_config(torch, dist, hp, [(zc, masks_c, mode_c)], C4_IMAGES * 256 * 256, args, world, dev,
                                                              flush, k_steps), scaling="strong", images_per_gpu=hi - lo)
        del zc, masks_c
        # configs[4]: the six 768-pixel tiles of one 2032 x 1344 image (inference_high_resolution.py:112-125), tile i -> rank i % world,
        # a rank's equal-shape tiles in one launch
        plan = cinf.tile_plan(*C5_IMAGE)
        mine = [i for i in range(len(plan)) if i % world == rank]
        groups = []
        for (th, tw), members in cinf.group_tiles([plan[i] for i in mine]).items():
            ids = [mine[k] for k in members]
            zs_, ms_ = [], [[], [], []]
            for i in ids:                                                            # tile i is "global image" 10000 + i of shape th x tw
                zi, mi, mode_c = hp.inputs(1, th, tw, 0.1, 0.8, 10000 + i)
                zs_.append(zi)
                for lvl in range(3):
                    ms_[lvl].append(mi[lvl])
            groups.append((torch.cat(zs_), tuple(torch.cat(v) for v in ms_), mode_c))
        cfg["c5_2032x1344_tiled_r0.1-0.8-0.1"] = dict(run_config(torch, dist, hp, groups, C5_IMAGE[0] * C5_IMAGE[1], args, world, dev, flush, k_steps),
                                                      scaling="strong", tiles=len(plan), tiles_this_rank=len(mine),
                                                      tile_shapes=sorted({(p[2], p[3]) for p in plan}, reverse=True))
        line["configs"] = cfg

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, cpu_sizes, _, check = cpu_run(args, 16, args.cpu_seconds, verify=True)
        # parity gate in the same run: indices and the five streams of every sample image, byte for byte
        idx_h = idx.cpu().view(B, -1)
        ind_h = ind.cpu().view(B, -1)
        for b, (c_ind, c_digest, c_dec) in enumerate(check):
            assert list(cpu_sizes[b]) == sizes_first[b].tolist(), f"CPU arm and GPU stream sizes differ on image {b}"
            assert torch.equal(c_ind, idx_h[b]), f"CPU arm and GPU indices differ on image {b}"
            assert c_digest == digests[b], f"CPU arm and GPU stream bytes differ on image {b}"
            assert torch.equal(c_dec, ind_h[b]), f"CPU arm and GPU decoded indices differ on image {b}"
        cpu["parity"] = (f"indices, decoded indices and the sha256 of all five streams of {len(check)} images identical to the GPU's "
                         "(bpp bit-exact)")
        line["cpu_baseline"] = cpu
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture (profiles/), filled in after
# each capture; a kernel without an entry reports null.
NCU_TRAFFIC_SOURCE = ("profiles/r2_kernels.txt: dram__bytes_read.sum + dram__bytes_write.sum of one launch on the 64-image batch under "
                      "ncu --set full (reads only: the 20 MB working set's writes stay in the 126 MB L2 within a launch); not re-measured in this run")
NCU_TRAFFIC = {"vq_warp_kernel": 6447104, "pack_kernel": 3525632, "unpack_decode_kernel": 249600, "unpack_assemble_kernel": 296704}


def _traffic(top, B):
    """The committed capture is of the 64-image batch: no figure for other batch sizes / kernels."""
    return NCU_TRAFFIC.get(top) if B == 64 else None


def algorithmic_step_bytes(B, h, w, stream_bytes):
    return algorithmic_bytes(B, h, w, stream_bytes)["step"]


def algorithmic_bytes(B, h, w, stream_bytes):
    """Algorithmic HBM bytes per launch of each kernel and of the whole step:
    per fine token: z 16 B in, z_q 16 B out, idx 8 B out (a1); masks 5.25 B in + streams out (pack);
    streams in + masks 5.25 B out (written as int64 like the reference) + ind 8 B out + quant 16 B out (unpack)."""
    n4 = B * h * w
    masks32 = n4 * 4 * (1 + 0.25 + 0.0625)
    masks64 = masks32            # SURVEY 8d counts the decoded masks at 5.25 B/token
    consts = 1024 * 16
    enc = n4 * (16 + 16 + 8) + masks32 + stream_bytes + consts
    dec = stream_bytes + masks64 + n4 * (8 + 16) + consts
    out = {
        "vq_fused_kernel": n4 * (16 + 16 + 8) + consts,
        "vq_warp_kernel": n4 * (16 + 16 + 8) + consts,       # the cell records it reads are not algorithmic
        "pack_kernel": n4 * 8 + masks32 + stream_bytes, "pack_chained_kernel": n4 * 8 + masks32 + stream_bytes,
        "unpack_decode_kernel": stream_bytes, "unpack_decode_chained_kernel": stream_bytes,
        "unpack_assemble_kernel": masks64 + n4 * (8 + 16) + consts,
        "encode_small_kernel": enc,                           # VQ + select + pack in one launch: the indices do not make a second trip
        "unpack_small_kernel": dec,                           # decode + re-assembly in one launch
    }
    out["step"] = enc + dec
    return out


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
