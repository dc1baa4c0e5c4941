"""Per-parameter gradient agreement (cosine, relative max-abs) of the generator backward vs torch CPU autograd."""
import sys, types, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch-tecogan_b200"))
import torch
from oracle import synth, tecogan_oracle as O
from tecogan_b200 import models
torch.set_num_threads(8)
for nres, shape, gain in [(2, (1, 51, 16, 16), 1.7), (2, (4, 51, 32, 32), 1.7), (16, (2, 51, 32, 32), 1.7), (16, (4, 51, 32, 32), 1.0)]:
    ref = O.OracleGenerator(3, nres)
    O.load_numpy_state(ref, synth.fill_state_dict(ref.state_dict(), seed=1, gain=gain))
    G = models.generator(3, types.SimpleNamespace(num_resblock=nres)); G.load_state_dict(ref.state_dict()); G = G.cuda()
    x = torch.from_numpy(synth.det_uniform(shape, 21, 0.0, 1.0)); n, _, h, w = shape
    target = torch.from_numpy(synth.det_uniform((n, 3, 4 * h, 4 * w), 22, 0.0, 1.0))
    ((ref(x) - target) ** 2).sum(dim=3).mean().backward()
    ((G(x.cuda()) - target.cuda()) ** 2).sum(dim=3).mean().backward()
    print("nres", nres, "shape", shape, "gain", gain)
    rows = []
    for (name, pr), (_, pg) in zip(ref.named_parameters(), G.named_parameters()):
        a, b = pg.grad.cpu().double().flatten(), pr.grad.double().flatten()
        rows.append((name, (a @ b / (a.norm() * b.norm() + 1e-30)).item(), (a - b).abs().max().item() / (b.abs().max().item() + 1e-30), (a.norm() / b.norm()).item()))
    for r in rows[:5] + rows[-8:]:
        print(f"  {r[0]:28s} cos {r[1]:.5f}  maxrel {r[2]:.4f}  norm ratio {r[3]:.4f}")
    print("  worst cos", min(r[1] for r in rows), "worst maxrel", max(r[2] for r in rows))
