"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on seeded
weights and seeded synthetic inputs.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py                 # every case (the BASELINE-scale ones take minutes each)
    python oracle/make_golden.py c2_grid30 ...   # selected cases

Only the *outputs* are stored (tiny); weights and inputs are re-created from their seeds by
cotracker_b200.synthetic on whichever machine runs the tests (the one real clip, BASELINE.json config 1's
assets/apple.mp4, travels as a 120x216 area-downsampled uint8 copy of its 50 decoded frames:
tests/golden/apple_frames_120x216.npz, made by `make_apple_fixture()` below).  Cases are listed in CASES below
and are the single source of truth for tests/test_golden*.py.  Predictor cases also store the model's visibility
(and confidence) probabilities -- captured with a forward hook on the unmodified reference model -- so a test
can tell a genuine visibility mismatch from a value sitting on the threshold.
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
    # T > 64: the multi-chunk time-attention path, odd T, time-embedding interpolation 60 -> 70
    "offline_long_T": dict(kind="model_offline", T=70, H=64, W=96, N=14, iters=2, wseed=21, vseed=22, qseed=23,
                           head_gain=5.0, vis_gain=30.0, window_len=60),
    # ---- predictor paths (reference predictor.py:70-98, :132-140, :161-164, :192-209, :255-264) --------------
    "pred_segm_mask": dict(kind="predictor_offline", T=4, H=160, W=224, grid=9, mask="box", iters=6, wseed=31,
                           vseed=32, head_gain=10.0, vis_gain=100.0, window_len=60),
    "pred_backward": dict(kind="predictor_offline", T=7, H=144, W=192, N=9, backward=True, iters=6, wseed=33,
                          vseed=34, qseed=35, head_gain=10.0, vis_gain=100.0, window_len=60),
    "pred_grid_query_frame": dict(kind="predictor_offline", T=8, H=144, W=192, grid=6, grid_query_frame=3,
                                  backward=True, iters=6, wseed=36, vseed=37, head_gain=10.0, vis_gain=100.0,
                                  window_len=60),
    "pred_dense": dict(kind="predictor_dense", T=3, H=48, W=160, iters=6, wseed=38, vseed=39, head_gain=10.0,
                       vis_gain=100.0, window_len=60),
    "pred_online_support_grid": dict(kind="predictor_online", T=24, H=160, W=224, N=6, add_support_grid=True,
                                     iters=6, wseed=40, vseed=41, qseed=42, head_gain=5.0, vis_gain=30.0,
                                     window_len=16),
    # ---- BASELINE.json configs at full size (reference CPU run: 16 s ... 2 min each) -------------------------
    # C1: assets/apple.mp4 (50 frames), grid_size=10, the demo.py:92-98 call pattern
    "c1_apple_grid10": dict(kind="predictor_offline", video="apple", grid=10, iters=6, wseed=1234, head_gain=1.0,
                            vis_gain=1.0, window_len=60),
    "c1_apple_grid10_stress": dict(kind="predictor_offline", video="apple", grid=10, iters=6, wseed=1234,
                                   head_gain=10.0, vis_gain=100.0, window_len=60),
    # C2: synthetic 512x512x16, grid_size=30
    "c2_grid30": dict(kind="predictor_offline", T=16, H=512, W=512, grid=30, iters=6, wseed=1234, vseed=0,
                      head_gain=1.0, vis_gain=1.0, window_len=60),
    "c2_grid30_stress": dict(kind="predictor_offline", T=16, H=512, W=512, grid=30, iters=6, wseed=1234, vseed=0,
                             head_gain=10.0, vis_gain=100.0, window_len=60),
    # headline: synthetic 512x512x16, grid_size=80 (N=6400) -- exactly bench.py's workload (same seeds)
    "headline_grid80": dict(kind="predictor_offline", T=16, H=512, W=512, grid=80, iters=6, wseed=1234, vseed=0,
                            head_gain=1.0, vis_gain=1.0, window_len=60),
    "headline_grid80_stress": dict(kind="predictor_offline", T=16, H=512, W=512, grid=80, iters=6, wseed=1234,
                                   vseed=0, head_gain=10.0, vis_gain=100.0, window_len=60),
    # C4: online predictor, 512x512 stream, window 16 / step 8, grid_size=50 (N=2500), 4 steps
    "c4_online_grid50": dict(kind="predictor_online", T=40, H=512, W=512, grid=50, iters=6, wseed=1234, vseed=0,
                             head_gain=5.0, vis_gain=30.0, window_len=16),
}
APPLE_FIXTURE = os.path.join(ROOT, "tests", "golden", "apple_frames_120x216.npz")


def make_apple_fixture():
    """Decode assets/apple.mp4 (BASELINE.json config 1) and store an area-downsampled uint8 copy (build box only)."""
    import cv2
    cap, frames = cv2.VideoCapture(os.path.join(REF, "assets", "apple.mp4")), []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(cv2.resize(cv2.cvtColor(f, cv2.COLOR_BGR2RGB), (216, 120), interpolation=cv2.INTER_AREA))
    np.savez_compressed(APPLE_FIXTURE, frames=np.stack(frames))


def box_mask(H, W):
    """[1,1,H,W] segmentation mask: an off-centre rectangle (keeps ~1/3 of a regular grid)."""
    m = torch.zeros(1, 1, H, W)
    m[:, :, H // 5: H // 5 * 4, W // 3: W // 8 * 7] = 1.0
    return m


def case_inputs(cfg):
    offline = cfg["kind"] in ("model_offline", "predictor_offline", "predictor_dense")
    sd = seeded_state_dict(cfg["wseed"], offline=offline, window_len=cfg["window_len"],
                           head_gain=cfg["head_gain"], vis_gain=cfg["vis_gain"])
    if cfg.get("video") == "apple":
        with np.load(APPLE_FIXTURE) as z:
            video = torch.from_numpy(z["frames"]).permute(0, 3, 1, 2)[None].float().contiguous()
        cfg = dict(cfg, T=video.shape[1], H=video.shape[3], W=video.shape[4])
    else:
        video = texture_video(cfg["T"], cfg["H"], cfg["W"], seed=cfg["vseed"])
    queries = None
    if "N" in cfg:
        queries = random_queries(cfg["N"], cfg["T"], cfg["H"], cfg["W"], seed=cfg["qseed"])
    return sd, video, queries


def predictor_kwargs(cfg, video, queries):
    """Keyword arguments of the public predictor call of a case (shared with tests/cases.py)."""
    kw = {}
    if cfg["kind"] == "predictor_dense":
        return kw
    if queries is not None:
        kw["queries"] = queries
    else:
        kw["grid_size"] = cfg["grid"]
    if cfg.get("grid_query_frame"):
        kw["grid_query_frame"] = cfg["grid_query_frame"]
    if cfg["kind"] == "predictor_online":
        if cfg.get("add_support_grid"):
            kw["add_support_grid"] = True
        return kw
    if cfg.get("mask") == "box":
        kw["segm_mask"] = box_mask(video.shape[3], video.shape[4]).to(video.device)
    if cfg.get("backward"):
        kw["backward_tracking"] = True
    return kw


def record_model_outputs(model):
    """Record (vis, conf) of every call of the unmodified reference model (the predictors call model.forward
    directly, so an instance-level wrapper rather than a forward hook)."""
    rec, fwd = [], model.forward

    def wrapped(*a, **k):
        o = fwd(*a, **k)
        rec.append((o[1].clone(), o[2].clone()))
        return o
    model.forward = wrapped
    return rec


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
        elif kind in ("predictor_offline", "predictor_dense"):
            p = CoTrackerPredictor(checkpoint=None, window_len=cfg["window_len"])
            p.model.load_state_dict(sd)
            probs = record_model_outputs(p.model)   # (vis, conf) probabilities of every model call, in call order
            tr, vi = p(video, **predictor_kwargs(cfg, video, queries))
            out = dict(tracks=tr, visibility=vi)
            if kind == "predictor_offline":
                out["prob_vis"] = probs[0][0]          # forward pass; support-grid columns still attached
                if cfg.get("backward"):
                    out["prob_vis_inv"] = probs[1][0].flip(1)
        elif kind == "predictor_online":
            p = CoTrackerOnlinePredictor(checkpoint=None, window_len=cfg["window_len"])
            p.model.load_state_dict(sd)
            probs = record_model_outputs(p.model)
            p(video_chunk=video, is_first_step=True, **predictor_kwargs(cfg, video, queries))
            k = 0
            for ind in range(0, video.shape[1] - p.step, p.step):
                tr, vi = p(video_chunk=video[:, ind:ind + p.step * 2],
                           add_support_grid=cfg.get("add_support_grid", False))
                out[f"tracks{k}"], out[f"visibility{k}"] = tr.clone(), vi.clone()
                out[f"prob_visconf{k}"] = probs[k][0] * probs[k][1]
                k += 1
        else:
            raise ValueError(kind)
    return {k: v.numpy() for k, v in out.items()}


def eval_case_inputs():
    """Seeded inputs of the EvaluationPredictor golden (tests/test_evaluation.py)."""
    sd = seeded_state_dict(51, offline=True, window_len=60, head_gain=10.0, vis_gain=100.0)
    video = texture_video(6, 160, 224, seed=52)
    queries = random_queries(5, 6, 160, 224, seed=53)
    return sd, video, queries


def make_eval_golden():
    """reference cotracker/models/evaluation_predictor.py:25-199, single-point (TAP-Vid protocol) and joint mode."""
    sys.path.insert(0, REF)
    from cotracker.models.build_cotracker import build_cotracker
    from cotracker.models.evaluation_predictor import EvaluationPredictor
    sd, video, queries = eval_case_inputs()
    m = build_cotracker(None, offline=True, window_len=60).eval()
    m.load_state_dict(sd)
    out = {}
    with torch.no_grad():
        for key, single in (("single", True), ("joint", False)):
            ev = EvaluationPredictor(m, single_point=single, grid_size=5, local_grid_size=8)
            tr, vi = ev(video, queries)
            out[f"tracks_{key}"], out[f"vis_{key}"] = tr.numpy(), vi.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "eval_predictor.npz"), **out)
    print("eval_predictor", {k: v.shape for k, v in out.items()})


def main():
    import time
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    if not os.path.exists(APPLE_FIXTURE):
        make_apple_fixture()
    names = sys.argv[1:] or (list(CASES) + ["eval_predictor"])
    if "eval_predictor" in names:
        names.remove("eval_predictor")
        make_eval_golden()
    for name in names:
        cfg = CASES[name]
        t0 = time.perf_counter()
        out = run_reference(cfg)
        print(f"{name}: reference ran {time.perf_counter() - t0:.1f} s")
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
