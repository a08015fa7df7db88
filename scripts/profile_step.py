"""One headline step (predictor call, N=6400, T=16, 512x512) inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off ... python scripts/profile_step.py [grid] [frames]
A number printed under ncu is never a bench value; this script prints nothing but 'done'."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cotracker_b200.predictor import CoTrackerPredictor
from cotracker_b200.synthetic import seeded_state_dict, texture_video

G = int(sys.argv[1]) if len(sys.argv) > 1 else 80
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = "cuda:0"
p = CoTrackerPredictor(checkpoint=None, window_len=60)
p.model.load_state_dict(seeded_state_dict(1234))
p = p.to(dev)
video = texture_video(T, 512, 512, seed=0).to(dev)
p(video, grid_size=G)            # warm-up (allocations, weight packing, cuDNN heuristics)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
p(video, grid_size=G)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
