"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on seeded
weights and seeded synthetic inputs.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Only the *outputs* are stored (tiny); weights and inputs are re-created from their seeds by
cotracker_b200.synthetic on whichever machine runs the tests.  Cases are listed in CASES below and are the
single source of truth for tests/test_golden*.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("COTRACKER_REFERENCE", "/root/reference")

from cotracker_b200.synthetic import random_queries, seeded_state_dict, texture_video  # noqa: E402

# name -> config.  kind: model_offline | model_online_stream | model_online_slide | predictor_offline | predictor_online
CASES = {
    # small feature maps (24x32 ... 3x4): every pyramid level hits the border-clamp path
    "offline_small": dict(kind="model_offline", T=6, H=96, W=128, N=20, iters=3, wseed=1234, vseed=1, qseed=2,
                          head_gain=1.0, vis_gain=1.0, window_len=60),
    # amplified heads: ~10-20 px of motion, vis/conf logits swing (stress regime of SURVEY.md 8c)
    "offline_stress": dict(kind="model_offline", T=8, H=128, W=160, N=24, iters=4, wseed=4321, vseed=3, qseed=4,
                           head_gain=10.0, vis_gain=100.0, window_len=60),
    # T == window_len: no time-embedding interpolation
    "offline_T_eq_window": dict(kind="model_offline", T=8, H=96, W=128, N=12, iters=2, wseed=99, vseed=5, qseed=6,
                                head_gain=3.0, vis_gain=10.0, window_len=8),
    # streaming online model: 3 chunks of a 32-frame video, queries entering in later windows
    "online_stream": dict(kind="model_online_stream", T=32, H=96, W=128, N=18, iters=3, wseed=77, vseed=7, qseed=8,
                          head_gain=5.0, vis_gain=30.0, window_len=16),
    # the online model sliding over a whole video in one call (is_online=False)
    "online_slide": dict(kind="model_online_slide", T=27, H=96, W=128, N=10, iters=2, wseed=78, vseed=9, qseed=10,
                         head_gain=5.0, vis_gain=30.0, window_len=16),
    # public predictor API at the model resolution (384x512 internally), regular grid
    "predictor_grid": dict(kind="predictor_offline", T=4, H=240, W=320, grid=5, iters=6, wseed=5, vseed=11,
                           head_gain=10.0, vis_gain=100.0, window_len=60),
    # public predictor API with explicit queries (adds the 6x6 support grid)
    "predictor_queries": dict(kind="predictor_offline", T=3, H=200, W=256, N=7, iters=6, wseed=6, vseed=12, qseed=13,
                              head_gain=10.0, vis_gain=100.0, window_len=60),
    # online predictor: first step + 2 steps of 16-frame chunks with stride 8
    "predictor_online": dict(kind="predictor_online", T=24, H=192, W=256, grid=4, iters=6, wseed=8, vseed=14,
                             head_gain=5.0, vis_gain=30.0, window_len=16),
}


def case_inputs(cfg):
    offline = cfg["kind"] in ("model_offline", "predictor_offline")
    sd = seeded_state_dict(cfg["wseed"], offline=offline, window_len=cfg["window_len"],
                           head_gain=cfg["head_gain"], vis_gain=cfg["vis_gain"])
    video = texture_video(cfg["T"], cfg["H"], cfg["W"], seed=cfg["vseed"])
    queries = None
    if "N" in cfg:
        queries = random_queries(cfg["N"], cfg["T"], cfg["H"], cfg["W"], seed=cfg["qseed"])
    return sd, video, queries


def run_reference(cfg):
    sys.path.insert(0, REF)
    from cotracker.models.build_cotracker import build_cotracker
    from cotracker.predictor import CoTrackerOnlinePredictor, CoTrackerPredictor

    sd, video, queries = case_inputs(cfg)
    kind = cfg["kind"]
    out = {}
    with torch.no_grad():
        if kind == "model_offline":
            m = build_cotracker(None, offline=True, window_len=cfg["window_len"]).eval()
            m.load_state_dict(sd)
            c, v, q, _ = m(video, queries, iters=cfg["iters"])
            out = dict(coords=c, vis=v, conf=q)
        elif kind == "model_online_slide":
            m = build_cotracker(None, offline=False, window_len=cfg["window_len"]).eval()
            m.load_state_dict(sd)
            c, v, q, _ = m(video, queries, iters=cfg["iters"], is_online=False)
            out = dict(coords=c, vis=v, conf=q)
        elif kind == "model_online_stream":
            m = build_cotracker(None, offline=False, window_len=cfg["window_len"]).eval()
            m.load_state_dict(sd)
            m.init_video_online_processing()
            S = cfg["window_len"]
            for k, ind in enumerate(range(0, cfg["T"] - S // 2, S // 2)):
                c, v, q, _ = m(video[:, ind:ind + S], queries, iters=cfg["iters"], is_online=True)
                out[f"coords{k}"], out[f"vis{k}"], out[f"conf{k}"] = c.clone(), v.clone(), q.clone()
        elif kind == "predictor_offline":
            p = CoTrackerPredictor(checkpoint=None, window_len=cfg["window_len"])
            p.model.load_state_dict(sd)
            if queries is None:
                tr, vi = p(video, grid_size=cfg["grid"])
            else:
                tr, vi = p(video, queries=queries)
            out = dict(tracks=tr, visibility=vi)
        elif kind == "predictor_online":
            p = CoTrackerOnlinePredictor(checkpoint=None, window_len=cfg["window_len"])
            p.model.load_state_dict(sd)
            p(video_chunk=video, is_first_step=True, grid_size=cfg["grid"])
            k = 0
            for ind in range(0, cfg["T"] - p.step, p.step):
                tr, vi = p(video_chunk=video[:, ind:ind + p.step * 2])
                out[f"tracks{k}"], out[f"visibility{k}"] = tr.clone(), vi.clone()
                k += 1
        else:
            raise ValueError(kind)
    return {k: v.numpy() for k, v in out.items()}


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, cfg in CASES.items():
        out = run_reference(cfg)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
