"""Evaluation harness -- SURVEY.md 8(f) rank 4.

    tapvid_metrics        -- reference cotracker/evaluation/core/eval_utils.py:12-138 (compute_tapvid_metrics)
    EvaluationPredictor   -- reference cotracker/models/evaluation_predictor.py:25-199

`tapvid_metrics` is plain numpy (no GPU): TAP-Vid occlusion accuracy, points-within-threshold (delta_avg) and
Jaccard (AJ) at 1/2/4/8/16 px.  `EvaluationPredictor` wraps a cotracker_b200 offline model the way the
reference's benchmark code drives it: one query point at a time with an 8x8 local grid and a 5x5 global grid as
helper tracks (single_point=True, the TAP-Vid protocol), or all queries jointly.  The model behind it is the same
CUDA path as everywhere else (libct3_b200.so); SIFT helper points (sift_size > 0) are not provided.
With no datasets or checkpoints in this environment the harness is exercised by scoring the B200 tracks against
the reference's tracks on synthetic clips (tests/test_evaluation.py): identical outputs score 1.0 everywhere.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

THRESHOLDS = (1, 2, 4, 8, 16)


def tapvid_metrics(query_points: np.ndarray, gt_occluded: np.ndarray, gt_tracks: np.ndarray,
                   pred_occluded: np.ndarray, pred_tracks: np.ndarray, query_mode: str) -> Dict[str, np.ndarray]:
    """TAP-Vid metrics per video.  Shapes: query_points [b,n,3] as (t, y, x); *_occluded [b,n,t] bool;
    *_tracks [b,n,t,2] as (x, y) in raster coordinates (the paper's numbers assume 256x256).
    query_mode "first": only frames strictly after the query frame are scored; "strided": every other frame."""
    b, n, t = gt_occluded.shape
    frames = np.arange(t)
    qf = np.round(query_points[..., 0]).astype(np.int32)                       # [b,n]
    if query_mode == "first":
        scored = frames[None, None, :] > qf[..., None]
    elif query_mode == "strided":
        scored = frames[None, None, :] != qf[..., None]
    else:
        raise ValueError("Unknown query mode " + query_mode)

    def per_video(mask):
        return np.sum(mask & scored, axis=(1, 2))

    out: Dict[str, np.ndarray] = {}
    # NB: the denominator is the number of scored points of the WHOLE batch (reference eval_utils.py:75-78)
    out["occlusion_accuracy"] = per_video(pred_occluded == gt_occluded) / np.sum(scored)
    gt_vis, pred_vis = ~gt_occluded.astype(bool), ~pred_occluded.astype(bool)
    d2 = np.sum(np.square(pred_tracks - gt_tracks), axis=-1)
    n_gt_vis = per_video(gt_vis)
    within_all, jac_all = [], []
    for thr in THRESHOLDS:
        close = d2 < thr * thr
        hit = close & gt_vis
        out[f"pts_within_{thr}"] = per_video(hit) / n_gt_vis
        # false positive: predicted visible where the ground truth is occluded or further than the threshold
        false_pos = per_video(pred_vis & (~gt_vis | ~close))
        out[f"jaccard_{thr}"] = per_video(hit & pred_vis) / (n_gt_vis + false_pos)
        within_all.append(out[f"pts_within_{thr}"])
        jac_all.append(out[f"jaccard_{thr}"])
    out["average_jaccard"] = np.mean(np.stack(jac_all, axis=1), axis=1)
    out["average_pts_within_thresh"] = np.mean(np.stack(within_all, axis=1), axis=1)
    return out


def points_on_a_grid(size: int, extent, center=None, device="cpu") -> torch.Tensor:
    """size x size grid of (x, y) points over an (H, W) extent around `center` (cy, cx), margin W/64, row-major
    (contract of reference model_utils.py:83-139)."""
    H, W = float(extent[0]), float(extent[1])
    if size == 1:
        return torch.tensor([W / 2, H / 2], device=device)[None, None]
    cy, cx = (H / 2, W / 2) if center is None else (float(center[0]), float(center[1]))
    m = W / 64
    ys = torch.linspace(m - H / 2 + cy, H / 2 + cy - m, size, device=device)
    xs = torch.linspace(m - W / 2 + cx, W / 2 + cx - m, size, device=device)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1).reshape(1, -1, 2)


class EvaluationPredictor(torch.nn.Module):
    """Benchmark-protocol wrapper around an offline CoTracker3 model (B = 1).

    forward(video [1,T,3,H,W] in 0..255, queries [1,N,3] = (t, x, y) in input pixels)
        -> (tracks [1,T,N,2] in input pixels, visibility*confidence [1,T,N] probabilities)
    """

    def __init__(self, cotracker_model, interp_shape: Tuple[int, int] = (384, 512), grid_size: int = 5,
                 local_grid_size: int = 8, single_point: bool = True, sift_size: int = 0,
                 num_uniformly_sampled_pts: int = 0, n_iters: int = 6, local_extent: int = 50) -> None:
        super().__init__()
        if sift_size > 0:
            raise NotImplementedError("SIFT helper points are not provided by the B200 build")
        self.grid_size = grid_size
        self.local_grid_size = local_grid_size
        self.single_point = single_point
        self.sift_size = 0
        self.interp_shape = interp_shape
        self.n_iters = n_iters
        self.num_uniformly_sampled_pts = num_uniformly_sampled_pts
        self.local_extent = local_extent
        self.model = cotracker_model
        self.model.eval()

    def _helpers(self, video, query: Optional[torch.Tensor]) -> torch.Tensor:
        """Helper tracks appended after the evaluated ones: [local grid around the query], global grid, random."""
        dev, extra = video.device, []
        if query is not None and self.local_grid_size > 0:
            loc = points_on_a_grid(self.local_grid_size, (self.local_extent, self.local_extent),
                                   (query[0, 0, 2].item(), query[0, 0, 1].item()), device=dev)
            extra.append(torch.cat([torch.zeros_like(loc[:, :, :1]), loc], dim=2))
        if self.grid_size > 0:
            g = points_on_a_grid(self.grid_size, video.shape[3:], device=dev)
            extra.append(torch.cat([torch.zeros_like(g[:, :, :1]), g], dim=2))
        if self.num_uniformly_sampled_pts > 0:
            k, T, (H, W) = self.num_uniformly_sampled_pts, video.shape[1], video.shape[3:]
            tt = torch.randint(0, T, (k, 1), device=dev).float()
            xy = torch.rand(k, 2, device=dev) * torch.tensor([W, H], device=dev, dtype=torch.float32)
            extra.append(torch.cat([tt, xy], dim=1)[None])
        return torch.cat(extra, dim=1) if extra else video.new_zeros(1, 0, 3)

    @torch.no_grad()
    def forward(self, video, queries):
        B, T, C, H, W = video.shape
        assert queries.shape[0] == 1 and queries.shape[2] == 3 and B == 1
        N = queries.shape[1]
        ih, iw = self.interp_shape
        video = F.interpolate(video.reshape(B * T, C, H, W), (ih, iw), mode="bilinear", align_corners=True)
        video = video.reshape(B, T, 3, ih, iw)
        queries = queries.clone()
        queries[:, :, 1] *= (iw - 1) / (W - 1)
        queries[:, :, 2] *= (ih - 1) / (H - 1)
        if self.single_point:
            tracks = video.new_zeros(B, T, N, 2)
            vis = video.new_zeros(B, T, N)
            conf = video.new_zeros(B, T, N)
            for i in range(N):
                q = queries[:, i:i + 1]
                q_all = torch.cat([q, self._helpers(video, q)], dim=1)
                tr, vi, cf, _ = self.model(video=video, queries=q_all, iters=self.n_iters)
                tracks[:, :, i], vis[:, :, i], conf[:, :, i] = tr[:, :, 0, :2], vi[:, :, 0], cf[:, :, 0]
        else:
            q_all = torch.cat([queries, self._helpers(video, None)], dim=1)
            tr, vi, cf, _ = self.model(video=video, queries=q_all, iters=self.n_iters)
            tracks, vis, conf = tr[:, :, :N].clone(), vi[:, :, :N], cf[:, :, :N]
        tracks[..., 0] *= (W - 1) / float(iw - 1)
        tracks[..., 1] *= (H - 1) / float(ih - 1)
        return tracks, vis * conf
