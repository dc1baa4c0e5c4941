"""Per-segment timeline of the frame kernel (tg_frame_set_trace): for every (layer, Cout chunk) segment the
time at which CTAs start their first item of it, relative to the kernel start.  Prints the span each segment
occupies (median start of next segment - median start of this one) next to its MMA-bound time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-tecogan_b200"))
from tecogan_b200 import _native as nt  # noqa: E402

lib = nt.lib()
N = int(os.environ.get("TG_N", "1"))
H, W = 180, 320
nres = 16
flat = (torch.rand(lib.tg_gen_param_count(nres), device="cuda") - 0.5) * 0.05
packed = torch.zeros(lib.tg_gen_packed_bytes(nres), dtype=torch.uint8, device="cuda")
nt.check(lib.tg_gen_pack(nt.ptr(flat), nres, nt.ptr(packed), nt.stream_ptr()))
ws = torch.empty(lib.tg_gen_workspace_bytes(N, H, W), dtype=torch.uint8, device="cuda")
x0 = torch.rand(N, H, W, 64, device="cuda").to(torch.bfloat16)
out = torch.empty(N, 3, 4 * H, 4 * W, device="cuda")
G = 148
NSEG = 44
trace = torch.zeros((NSEG + 1 + 32) * G, dtype=torch.int64, device="cuda")


def run():
    nt.check(lib.tg_gen_forward(nt.ptr(packed), nres, nt.ptr(x0), nt.ptr(out), None, nt.ptr(ws), ws.numel(), N, H, W, 2, nt.stream_ptr()))


for _ in range(3):
    run()
torch.cuda.synchronize()
nt.check(lib.tg_frame_set_trace(nt.ptr(trace), trace.numel() * 8))
run()
torch.cuda.synchronize()
nt.check(lib.tg_frame_set_trace(None, 0))
stats = trace[(NSEG + 1) * G:].view(G, 32).cpu().double()
t = trace[:(NSEG + 1) * G].view(NSEG + 1, G).cpu().double()
t0 = t[t > 0].min()
names = ["conv.0"] + [f"res{i // 2}.{'0' if i % 2 == 0 else '2'}" for i in range(32)] + ["convT64", "ct2.0 64@2x", "ct2.2 64@2x",
         "ct3.0 64->128 c0", "ct3.0 64->128 c1", "ct3.2 128->128 c0", "ct3.2 128->128 c1", "convT128 c0", "convT128 c1",
         "ct6 128->64@4x", "out 64->3", "END"]
# MMA-bound time per segment, in units of (36 MMAs x 75 cycles): tall geometry items/148 * 36*kchunks MMAs * 75 cycles
# (N=64); wide geometry 12*kchunks MMAs * 139 cycles (N=192) per 30x4 tile; / 1.965 GHz
px = N * H * W
WIDE = os.environ.get("TG_FRAME_WIDE", "1") != "0"
if WIDE:
    wt = lambda s: N * (-(-W * s // 30)) * (-(-H * s // 4)) * (12 * 139) / (36 * 75)
    tall = lambda s: N * (-(-W * s // 8)) * (-(-H * s // 16))
    bound = [wt(1)] * 33 + [tall(1)] + [wt(2)] * 4 + [wt(2) * 2] * 2 + [tall(2) * 2] * 2 + [wt(4) * 2, wt(4) * 67 / 139]
else:
    tiles1, tiles2, tiles4 = N * 40 * 12, N * 80 * 23, N * 160 * 45
    bound = [tiles1] * 34 + [tiles2] * 4 + [tiles2 * 2] * 4 + [tiles4 * 2, tiles4 * 52.7 / 75]
prev = None
print(f"N={N}  total {(t[NSEG].max() - t0) / 1e3:.1f} us")
for s in range(NSEG + 1):
    row = t[s]
    row = row[row > 0]
    if row.numel() == 0:
        continue
    med, lo, hi = row.median().item() - t0.item(), row.min().item() - t0.item(), row.max().item() - t0.item()
    if prev is not None:
        b = bound[prev[0]] / 148 * 36 * 75 / 1.965
        print(f"{names[prev[0]]:22s} start med {prev[1] / 1e3:8.1f} us (min {prev[2] / 1e3:7.1f} max {prev[3] / 1e3:7.1f})  span {(med - prev[1]) / 1e3:7.1f} us   mma-bound {b / 1e3:6.1f} us")
    prev = (s, med, lo, hi)

# stall accounting: mean cycles per CTA spent in each wait (leader CTAs only for the issuer in pair mode)
lab = ["mma: wait acc drained", "mma: wait weights", "mma: wait A stage", "mma: loop total",
       "epi(w2): wait acc ready", "epi(w2): math+store (+ld if not wide bf16)", "epi(w2): publish handoff", "epi(w2): loop total", "epi(w2): tmem ld (wide bf16)",
       "prod: wait weight slot", "prod: wait dependencies", "prod: wait A stage free", "prod: loop total",
       "pub: wait tile stored", "pub: release + arrive", "-", "-", "epi(w2): accumulator hand-back", "-", "-",
       "mma: issue block (MMAs + commits)", "-", "prod: issue block (fence + expect_tx + TMA)"]
for k, name in enumerate(lab):
    col = stats[:, k]
    col = col[col > 0]
    if name != "-" and col.numel():
        print(f"stat {name:28s} mean {col.mean().item() / 1e3:9.1f} kcycles  (max {col.max().item() / 1e3:9.1f}, {col.numel()} CTAs)")
