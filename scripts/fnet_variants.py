"""Encoder (fnet) options: timing and feature error of cuDNN benchmark mode / TF32 / channels_last vs plain fp32."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cotracker_b200.build import build_cotracker
from cotracker_b200.synthetic import seeded_state_dict, texture_video

dev = "cuda:0"
m = build_cotracker(None, offline=True, window_len=60)
m.load_state_dict(seeded_state_dict(1234)); m = m.to(dev).eval()
x = (2 * (texture_video(16, 384, 512, seed=0)[0] / 255) - 1).to(dev)

def run(tf32, bench, cl):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = bench
    net = m.fnet.to(memory_format=torch.channels_last) if cl else m.fnet.to(memory_format=torch.contiguous_format)
    inp = x.contiguous(memory_format=torch.channels_last) if cl else x.contiguous()
    with torch.no_grad():
        for _ in range(3): y = net(inp)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): y = net(inp)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 5 * 1e3, y.float().contiguous()

base_ms, base = run(False, False, False)
print(f"fp32 plain          {base_ms:6.2f} ms")
for name, args in [("fp32 benchmark", (False, True, False)), ("fp32 channels_last", (False, False, True)),
                   ("fp32 bench+cl", (False, True, True)), ("tf32", (True, False, False)), ("tf32 bench+cl", (True, True, True))]:
    ms, y = run(*args)
    err = float((y - base).abs().max()); rel = err / float(base.abs().max())
    print(f"{name:20s}{ms:6.2f} ms   max|d feat| {err:.3e} (rel {rel:.2e})", flush=True)
