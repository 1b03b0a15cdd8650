import sys; sys.path.insert(0, ".")
import numpy as np, torch, cgic_b200 as cg
from oracle import oracle as orc
def truth(x, p):
    x = x.astype(np.float64)
    g = 0.2989 * x[:, 0] + 0.5870 * x[:, 1] + 0.1140 * x[:, 2]
    B, H, W = g.shape
    pt = g.reshape(B, H // p, p, W // p, p).transpose(0, 1, 3, 2, 4).reshape(B, H // p, W // p, p * p)
    bins = np.linspace(-1, 1, 32)
    k = np.exp(-0.5 * ((pt[..., None] - bins) / 0.01) ** 2).mean(-2)
    pdf = k / (k.sum(-1, keepdims=True) + 1e-40) + 1e-40
    return -(pdf * np.log(pdf)).sum(-1)
g = torch.Generator().manual_seed(3)
for name, x in (("rand", torch.rand(2, 3, 128, 160, generator=g)), ("smooth", torch.rand(2, 3, 8, 10, generator=g).repeat_interleave(16, -1).repeat_interleave(16, -2) * 0.9 + 0.05 * torch.rand(2, 3, 128, 160, generator=g)), ("flat", torch.full((1, 3, 64, 64), 0.37))):
    e8, e16 = cg.entropy_pair(x.cuda())
    for p, e in ((8, e8), (16, e16)):
        o = orc.entropy(x.numpy(), p); t = truth(x.numpy(), p)
        rel = lambda a: (np.abs(a - t) / np.maximum(np.abs(t), 1e-9)).max()
        print(name, p, "GPU vs fp64 %.2e   oracle(fp32, reference order) vs fp64 %.2e   GPU vs oracle %.2e" % (rel(e.cpu().numpy()), rel(o), (np.abs(e.cpu().numpy() - o) / np.maximum(np.abs(o), 1e-9)).max()), "min entropy %.3g" % t.min())
