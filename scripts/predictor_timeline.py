"""Where does one predictor call spend its time?  CUDA-event segments around the engine calls + wall clock."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cotracker_b200 import engine
from cotracker_b200.predictor import CoTrackerPredictor
from cotracker_b200.synthetic import seeded_state_dict, texture_video

G = int(sys.argv[1]) if len(sys.argv) > 1 else 80
dev = "cuda:0"
p = CoTrackerPredictor(checkpoint=None, window_len=60)
p.model.load_state_dict(seeded_state_dict(1234))
p = p.to(dev)
video = texture_video(16, 512, 512, seed=0).to(dev)
marks = []

def wrap(obj, name, label):
    fn = getattr(obj, name)
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); r = fn(*a, **k); e1.record(); t1 = time.perf_counter()
        marks.append((label, e0, e1, (t1 - t0) * 1e3)); return r
    setattr(obj, name, w)

wrap(p.model, "_encode", "fnet")
wrap(engine, "prepare_pyramid", "pyramid")
wrap(engine, "sample_support", "support")
wrap(engine, "update_loop", "update_loop")
for rep in range(4):
    marks.clear()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s0.record(); p(video, grid_size=G); s1.record(); t_issue = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize(); t_all = (time.perf_counter() - t0) * 1e3
    seg = " ".join(f"{l}: gpu {a.elapsed_time(b):.1f} / cpu {c:.1f}" for l, a, b, c in marks)
    first = marks[0][1]
    print(f"rep{rep}: total gpu {s0.elapsed_time(s1):.1f} ms, wall {t_all:.1f} ms, cpu-issue {t_issue:.1f} ms | pre {s0.elapsed_time(first):.1f} | {seg} | post {marks[-1][2].elapsed_time(s1):.1f}", flush=True)
