"""GPU box: bisect the CUDA-graph capture failure of the training step."""
import os, sys, traceback, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch-tecogan_b200"))
import torch
from oracle import synth, train_oracle as TO
from tecogan_b200 import models, optim, train as T

variant = sys.argv[1]
args = TO.default_train_args(num_resblock=2, discrim_resblocks=1, discrim_channels=64, crop_size=32)
torch.manual_seed(1)
G, D = models.generator(3, args).cuda(), models.discriminator(args).cuda()
og = torch.optim.Adam(G.parameters(), 1e-4); od = torch.optim.Adam(D.parameters(), 1e-4)
r_in = torch.rand(2, 10, 3, 32, 32, device="cuda"); r_tg = torch.rand(2, 10, 3, 128, 128, device="cuda")
s_in, s_tg = torch.empty_like(r_in), torch.empty_like(r_tg)
dt = torch.ones((), device="cuda")
if variant in ("F", "G", "H"):
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        for i in range(3):
            if variant == "F":
                T.FRVSR_Train(r_in, r_tg, args, D, G, i, 0., 0., og, od)
            elif variant == "G":
                out = T.FRVSR_Train(r_in, r_tg, args, D, G, i, 0., 0., og, od)
            else:
                out = T.FRVSR_Train(r_in, r_tg, args, D, G, i, 0., 0., og, od)
                out = None
        msgs = [str(x.message)[:100] for x in w if "capture" in str(x.message)]
    gs = list(T._graphs.values())[-1]
    print(f"variant {variant}: captured={gs.graph is not None} failed={gs.failed} {msgs}")
    sys.exit(0)
for i in range(2):
    T.TecoGAN(r_in, r_tg, D, G, args, i, 0., 0., og, od)
torch.cuda.synchronize()
try:
    if variant == "E":
        gs = T._GraphedStep(r_in, r_tg)
        out = gs.run(r_in, r_tg, args, D, G, 2, 0., 0., og, od)
    else:
        a, b = r_in, r_tg
        if variant in ("B", "D"):
            a, b = s_in, s_tg
            s_in.copy_(r_in); s_tg.copy_(r_tg)
        if variant in ("C", "D"):
            for net, opt in ((G, og), (D, od)):
                optim.FlatAdam.adopt(net, opt).refresh_lr()
        if variant == "D":
            dt.fill_(1.0)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            out = T.TecoGAN(a, b, D, G, args, 2, 0., 0., og, od, _dt_ratio_dev=dt)
        g.replay()
    torch.cuda.synchronize()
    print(f"variant {variant}: capture OK gen_loss {float(out.gen_loss.detach()):.4f}")
except Exception as e:
    print(f"variant {variant}: FAILED {type(e).__name__} {str(e)[:120]}")
