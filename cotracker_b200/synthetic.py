"""Deterministic synthetic inputs and seeded weights (no network: no datasets, no released checkpoints).

Everything here is reproducible bit-for-bit from a seed with the CPU generator, so the same tensors can be
rebuilt on the GPU box without shipping 100 MB state dicts: tests, smoke() and bench.py all use these.
"""
from __future__ import annotations

import torch

from .build import build_cotracker


def seeded_state_dict(seed: int = 1234, offline: bool = True, window_len: int = 60, head_gain: float = 1.0,
                      vis_gain: float = 1.0):
    """Random-init CoTracker3 weights.  `head_gain` / `vis_gain` scale flow_head / vis_conf_head to stress the
    feedback loop (x10 / x100 gives ~20 px of motion, closer to trained weights)."""
    torch.manual_seed(seed)
    model = build_cotracker(None, offline=offline, window_len=window_len)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    if head_gain != 1.0:
        sd["updateformer.flow_head.weight"] *= head_gain
    if vis_gain != 1.0:
        sd["updateformer.vis_conf_head.weight"] *= vis_gain
    return sd


def texture_video(T: int, H: int, W: int, seed: int = 0, cell: int = 8, shift=(1, 2)) -> torch.Tensor:
    """[1,T,3,H,W] float32 with integer values in 0..255: a random low-resolution texture, nearest-upsampled
    by `cell` and translated by `shift` pixels (y,x) per frame.  Integer-only construction => bit-identical on
    every machine; coherent motion => tracks actually move."""
    g = torch.Generator().manual_seed(seed)
    hh, ww = (H + cell - 1) // cell + 1, (W + cell - 1) // cell + 1
    base = torch.randint(0, 256, (3, hh, ww), generator=g).float()
    big = base.repeat_interleave(cell, dim=1).repeat_interleave(cell, dim=2)
    frames = []
    for t in range(T):
        f = torch.roll(big, shifts=(t * shift[0], t * shift[1]), dims=(1, 2))
        frames.append(f[:, :H, :W])
    return torch.stack(frames)[None].contiguous()


def noise_video(T: int, H: int, W: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (1, T, 3, H, W), generator=g).float()


def random_queries(N: int, T: int, H: int, W: int, seed: int = 0, first_frame_only: bool = False) -> torch.Tensor:
    """[1,N,3] (t,x,y) queries; some are placed near / on the border to exercise the clamp path."""
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(N) if first_frame_only else torch.randint(0, T, (N,), generator=g).float()
    x = torch.rand(N, generator=g) * (W - 1)
    y = torch.rand(N, generator=g) * (H - 1)
    if N >= 4:
        x[0], y[0] = 0.0, 0.0
        x[1], y[1] = float(W - 1), float(H - 1)
        x[2], y[2] = 1.5, float(H - 1) - 0.25
    return torch.stack([t, x, y], dim=-1)[None]
