"""Soak test of the correlation kernel's dynamic unit queue: many launches with random track counts (fewer units than
SMs, exactly as many, a few more, many more), odd frame counts and samples outside the maps, each checked against the
exact-fp32 SIMT kernel.  Run under `timeout`; a protocol bug shows up as a trap (CUDA error), never as a hang.
    timeout 300 python scripts/corr_soak.py [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cotracker_b200 import engine

n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = "cuda:0"
g = torch.Generator().manual_seed(7)
H4, W4 = 64, 72
worst = 0.0
for it in range(n_launch):
    T = int(torch.randint(1, 21, (1,), generator=g))
    N = int([1, 2, 36, 37, 38, 40, 74, 148, 149, 300, 777][int(torch.randint(0, 11, (1,), generator=g))])
    fmaps = torch.randn(T, 128, H4, W4, generator=g).to(dev)
    pyr = engine.prepare_pyramid(fmaps)
    support = torch.randn(4, 49, N, 128, generator=g)
    support = (support / support.norm(dim=-1, keepdim=True)).to(dev)
    coords = (torch.rand(T, N, 2, generator=g) * torch.tensor([W4 + 8.0, H4 + 8.0]) - 4.0).to(dev)
    valid = (torch.rand(N, generator=g) > 0.1).to(torch.uint8).to(dev)
    engine.set_option("corr", 0)
    got = engine.corr_sample(pyr, H4, W4, support, valid, coords)
    if it % 10 == 0:
        engine.set_option("corr", 1)
        want = engine.corr_sample(pyr, H4, W4, support, valid, coords)
        engine.set_option("corr", 0)
        err = float((got - want).abs().max())
        worst = max(worst, err)
        assert err < 5e-4, (it, T, N, err)
torch.cuda.synchronize()
print(f"corr_soak: {n_launch} launches ok, worst |product kernel - exact fp32 kernel| = {worst:.2e}")
