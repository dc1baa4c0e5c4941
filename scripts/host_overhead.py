"""Host-side cost of enqueueing the frame loop: time for tg_gen_clip_forward (100 frames) to RETURN (asynchronous)
against the time the GPU needs for it.  The loop stays GPU-bound as long as the first is well below the second."""
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-tecogan_b200"))
from tecogan_b200 import models  # noqa: E402
from tecogan_b200.pipeline import ClipPipeline  # noqa: E402

B, T, H, W = int(os.environ.get("TG_B", "2")), 100, 180, 320
G = models.generator(3, types.SimpleNamespace(num_resblock=16)).cuda().eval()
pipe = ClipPipeline(G, B, T, H, W)
lr = torch.rand((B, T, 3, H, W), device="cuda") * 0.25
out = torch.empty((B, T, 3, 4 * H, 4 * W), device="cuda")
for _ in range(2):
    pipe.run_device(lr, out)
torch.cuda.synchronize()
t0 = time.perf_counter()
pipe.run_device(lr, out)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"clips {B}: enqueue of {T} frames returned after {(t1 - t0) * 1e3:.1f} ms ({(t1 - t0) / T * 1e6:.0f} us per frame), "
      f"GPU finished after {(t2 - t0) * 1e3:.1f} ms ({(t2 - t0) / T * 1e6:.0f} us per frame)")
