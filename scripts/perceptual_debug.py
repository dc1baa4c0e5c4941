import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch-tecogan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import synth
from tecogan_b200 import perceptual
import test_gpu_perceptual as TP
P = perceptual.PerceptualStandIn("cuda", seed=19)
gen = torch.from_numpy(synth.det_uniform((2, 3, 32, 48), 5, 0.05, 0.95)).cuda().requires_grad_(True)
tgt = torch.from_numpy(synth.det_uniform((2, 3, 32, 48), 6, 0.05, 0.95)).cuda()
f = P.features(torch.cat((gen, tgt), 0))
for k, v in f.items():
    for i, t in enumerate(v):
        t.retain_grad()
        print(k, i, tuple(t.shape), "nan", int(torch.isnan(t.float()).sum()), "absmax", float(t.float().abs().max()), "zero-frac", float((t == 0).float().mean()))
loss, per = P.loss(gen, tgt)
print("loss", float(loss), [float(x) for x in per])
loss.backward()
print("ours grad nan", int(torch.isnan(gen.grad).sum()), "absmax", float(gen.grad.nan_to_num().abs().max()))
gen.grad = None
want = TP._torch_loss(P, gen, tgt); want.backward()
print("torch loss", float(want), "grad nan", int(torch.isnan(gen.grad).sum()), "absmax", float(gen.grad.nan_to_num().abs().max()))
