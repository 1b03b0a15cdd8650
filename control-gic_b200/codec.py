"""HuffmanCoding / BinaryCoding -- drop-ins for CGIC/tools/indices_coding.py and mask_coding.py.

Same construction (`HuffmanCoding(frequency)` with the model's `embedding_counter` mapping,
`BinaryCoding()`), same attributes (`codes`, `reverse_mapping`) and the two methods the model
calls: `compress(tensor, path) -> path` and `decompress_string(path) -> list | None`, producing
and consuming byte-identical files.  The table is built by the C-ABI's heapq-exact host builder
(one device->host transfer of all counters instead of 1024 `.item()` syncs); packing and
unpacking run on the GPU.  `.table` exposes the native handle that the fused batched
pack / unpack kernels (CGIC.compress) use.
"""
from __future__ import annotations

import torch

from . import ops


def _default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("the B200 codec needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class HuffmanCoding:
    def __init__(self, frequency):
        keys = [int(k) for k in frequency.keys()]                      # iteration order = heap push order
        vals = list(frequency.values())
        if len(vals) and isinstance(vals[0], torch.Tensor):
            flat = torch.cat([v.detach().reshape(-1)[:1] for v in vals]).cpu()
            counts = flat.to(torch.float64).trunc().to(torch.int64).tolist()          # int(value.item())
        else:
            counts = [int(v) for v in vals]
        K = len(keys)
        if sorted(keys) != list(range(K)):
            raise ValueError("frequency keys must be the symbols 0..K-1")
        freq = [0] * K
        for k, c in zip(keys, counts):
            freq[k] = c
        self.table = ops.HuffTable(freq, keys)
        self.heap = []
        self.codes = self.table.codes()
        self.reverse_mapping = {c: s for s, c in self.codes.items()}

    def compress(self, info, output_path):
        t = info if isinstance(info, torch.Tensor) else torch.as_tensor(info)
        if not t.is_cuda:
            t = t.to(_default_device())
        data = ops.huff_encode(t, self.table)
        with open(output_path, "wb") as f:
            f.write(data)
        return output_path

    def decompress_string(self, path, device=None):
        with open(path, "rb") as f:
            data = f.read()
        return ops.huff_decode(data, self.table, device or _default_device())


class BinaryCoding:
    def __init__(self):
        self.heap = []
        self.codes = {0: "0", 1: "1"}
        self.reverse_mapping = {"0": 0, "1": 1}

    def compress(self, info, output_path):
        t = info if isinstance(info, torch.Tensor) else torch.as_tensor(info)
        if not t.is_cuda:
            t = t.to(_default_device())
        data = ops.bits_encode(t)
        with open(output_path, "wb") as f:
            f.write(data)
        return output_path

    def decompress_string(self, path, device=None):
        with open(path, "rb") as f:
            data = f.read()
        return ops.bits_decode(data, device or _default_device())
