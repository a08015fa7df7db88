"""CNN feature encoder feeding the hot loop.

SURVEY.md §8(f) rank 1 ("next" row): the encoder (BasicEncoder, reference blocks.py:141-219) stays on stock
PyTorch/cuDNN in this round; it only has to (a) hold the `fnet.*` checkpoint keys and (b) produce the same
[T,128,H/4,W/4] feature maps in fp32 (TF32 convolutions are disabled for the call: single-pass TF32 breaks
the 1e-3 px parity budget).

Architecture (as specified by the checkpoint layout): 7x7/2 conv(3->64) + IN + ReLU; four stages of two
residual units (64/1, 96/2, 128/2, 128/2) with InstanceNorm; the four stage outputs are bilinearly resized
(align_corners=True) to H/4 x W/4, concatenated (416 ch) -> 3x3 conv(256) + IN + ReLU -> 1x1 conv(128).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _ResidualUnit(nn.Module):
    """conv3x3 -> IN -> ReLU -> conv3x3 -> IN -> ReLU, plus identity (or 1x1-conv+IN when striding)."""

    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.stride = stride
        if stride != 1:
            # index 0 keeps the checkpoint key `downsample.0.{weight,bias}`; the norm has no parameters
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride))
        else:
            self.downsample = None

    def forward(self, x):
        y = F.relu(F.instance_norm(self.conv1(x)))
        y = F.relu(F.instance_norm(self.conv2(y)))
        if self.downsample is not None:
            x = F.instance_norm(self.downsample(x))
        return F.relu(x + y)


class BasicEncoder(nn.Module):
    def __init__(self, input_dim: int = 3, output_dim: int = 128, stride: int = 4):
        super().__init__()
        self.stride = stride
        widths = [output_dim // 2, output_dim // 4 * 3, output_dim, output_dim]
        self.conv1 = nn.Conv2d(input_dim, widths[0], 7, stride=2, padding=3)
        cin = widths[0]
        for i, (w, s) in enumerate(zip(widths, [1, 2, 2, 2]), start=1):
            setattr(self, f"layer{i}", nn.Sequential(_ResidualUnit(cin, w, s), _ResidualUnit(w, w, 1)))
            cin = w
        self.conv2 = nn.Conv2d(sum(widths), output_dim * 2, 3, padding=1)
        self.conv3 = nn.Conv2d(output_dim * 2, output_dim, 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def stages(self, x):
        """Stem + four residual stages -> their 4 outputs (64@1/2, 96@1/4, 128@1/8, 128@1/16 of the input)."""
        x = F.relu(F.instance_norm(self.conv1(x)))
        feats = []
        for i in range(1, 5):
            x = getattr(self, f"layer{i}")(x)
            feats.append(x)
        return feats

    def forward_front(self, x):
        """Stages + bilinear resize (align_corners) + concat -> [T, 416, H/4, W/4] (the input of conv2), PyTorch."""
        H, W = x.shape[-2:]
        size = (H // self.stride, W // self.stride)
        return torch.cat([F.interpolate(f, size, mode="bilinear", align_corners=True) for f in self.stages(x)], dim=1)

    def forward(self, x):
        x = F.relu(F.instance_norm(self.conv2(self.forward_front(x))))
        return self.conv3(x)
