"""Per-stage timing of one predictor call at growing grid sizes (progressively flushed, safe under `timeout`)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cotracker_b200 import engine
from cotracker_b200.predictor import CoTrackerPredictor, get_points_on_a_grid
from cotracker_b200.synthetic import seeded_state_dict, texture_video

dev = "cuda:0"
engine.set_option("corr", int(os.environ.get("CT3_CORR", "0")))   # 0 product, 1 SIMT, 2 sample-then-correlate
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
grids = [int(g) for g in sys.argv[1].split(",")] if len(sys.argv) > 1 else [10, 20, 40, 80]
p = CoTrackerPredictor(checkpoint=None, window_len=60)
p.model.load_state_dict(seeded_state_dict(1234))
p = p.to(dev)
m = p.model
video = texture_video(T, 384, 512, seed=0).to(dev)

def sync_time(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3

frames = 2.0 * (video[0] / 255.0) - 1.0
pyr, ms = sync_time(lambda: m._encode(frames, 200)); print(f"encoder+pyramid {ms:.1f} ms", flush=True)
pyr, ms = sync_time(lambda: m._encode(frames, 200)); print(f"encoder+pyramid(2) {ms:.1f} ms", flush=True)
for G in grids:
    N = G * G
    pts = get_points_on_a_grid(G, (384, 512), device=dev)[0]
    qf = torch.zeros(N, dtype=torch.int32, device=dev)
    qc = (pts / 4).contiguous()
    sup, ms = sync_time(lambda: engine.sample_support(pyr, T, 96, 128, qf, qc)); print(f"G={G} support {ms:.2f} ms", flush=True)
    coords = qc[None].expand(T, N, 2).contiguous(); vis = torch.zeros(T, N, device=dev); conf = torch.zeros(T, N, device=dev)
    ws = torch.empty(engine.workspace_bytes(T, N, 96, 128), dtype=torch.uint8, device=dev)
    packed = m.packed_weights(torch.device(dev)); te = m.interpolate_time_embed(T).to(dev)
    for rep in range(2):
        engine.profile_enable(True)
        _, ms = sync_time(lambda: engine.update_loop(packed, pyr, 96, 128, sup, None, coords, vis, conf, te, 1, ws))
        cat_ms, cat_n, fl = engine.profile_read(); engine.profile_enable(False)
        print(f"G={G} N={N} iter {ms:.1f} ms | " + " ".join(f"{k}={v:.2f}ms/{cat_n[k]}" for k, v in cat_ms.items()) +
              f" | gemm {fl / max(cat_ms['gemm'], 1e-9) / 1e9:.1f} TFLOP/s", flush=True)
    del ws
    _, ms = sync_time(lambda: p(video[None][0:1] if video.dim() == 4 else video, grid_size=G)); print(f"G={G} predictor call {ms:.1f} ms", flush=True)
