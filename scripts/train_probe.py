"""Training-step probe: a few FRVSR_Train steps of cfg4 (default) or cfg5 (TG_CFG=5) for ncu / timing."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-tecogan_b200"))
import bench  # noqa: E402
from tecogan_b200 import _native as _nt  # noqa: E402
if os.environ.get("TG_LIB"):                               # measurement only: another build of the library (same-box A/B)
    _nt.LIB_PATH = os.path.abspath(os.environ["TG_LIB"])
from tecogan_b200 import models, train as T  # noqa: E402

cfg5 = os.environ.get("TG_CFG", "4") == "5"
crop, b = (64, int(os.environ.get("TG_B", "32"))) if cfg5 else (32, int(os.environ.get("TG_B", "4")))
steps, warm = int(os.environ.get("TG_STEPS", "5")), int(os.environ.get("TG_WARM", "3"))
dev = torch.device("cuda", 0)
args = bench.train_args(crop)
torch.manual_seed(1)
G, D = models.generator(3, args).to(dev), models.discriminator(args).to(dev)
og = torch.optim.Adam(G.parameters(), 1e-4, betas=(0.9, 0.999), eps=1e-8)
od = torch.optim.Adam(D.parameters(), 1e-4, betas=(0.9, 0.999), eps=1e-8)
r_in = torch.rand((b, 10, 3, crop, crop), device=dev)
r_tg = torch.rand((b, 10, 3, 4 * crop, 4 * crop), device=dev)
for i in range(warm):
    T.FRVSR_Train(r_in, r_tg, args, D, G, i, 0.0, 0.0, og, od)
torch.cuda.synchronize()
if os.environ.get("TG_STAGES"):
    # coarse stage timing with events around the pieces of one step
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []
    def mark(name):
        e = ev(); e.record(); marks.append((name, e))
    mark("start")
    gen_tb = G.forward_clip_train(r_in); mark("G forward (10 frames)")
    real_in, fake_in = T.discriminator_inputs(r_in, r_tg, gen_tb, args); mark("D input assembly x2")
    pr, rl = D(real_in); pf, fl = D(fake_in); mark("D forward x2")
    content = torch.mean(torch.sum(torch.square(gen_tb.transpose(0, 1) - r_tg), dim=[4])); mark("content loss")
    content.backward(); mark("G backward (batched)")
    dl = torch.mean(-(torch.log(1 - pf + 1e-12) + torch.log(pr + 1e-12))); dl.backward(); mark("D backward x2")
    og.step(); od.step(); mark("Adam x2")
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        print(f"{n1:28s} {e0.elapsed_time(e1):8.3f} ms")
t0 = time.perf_counter()
for i in range(steps):
    T.FRVSR_Train(r_in, r_tg, args, D, G, warm + i, 0.0, 0.0, og, od)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
print(f"cfg{'5' if cfg5 else '4'} b={b}: {dt * 1e3:.2f} ms/step, {b / dt:.1f} clips/s, {bench.train_step_flops(b, 10, crop) / dt / 1e12:.1f} TFLOP/s")
