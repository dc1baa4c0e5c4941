import sys, types, math
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/pytorch-tecogan_b200')
import torch
from oracle import synth, tecogan_oracle as O
from tecogan_b200 import models
torch.set_num_threads(8)
for (nb,ch,n) in [(4,128,3),(4,128,12)]:
    ref = O.OracleDiscriminator(nb, ch, 48)
    O.load_numpy_state(ref, synth.fill_state_dict(ref.state_dict(), seed=2, gain=1.0))
    D = models.discriminator(types.SimpleNamespace(discrim_resblocks=nb, discrim_channels=ch, crop_size=32))
    D.load_state_dict(ref.state_dict()); D=D.cuda(); ref.train(); D.train()
    x = torch.from_numpy(synth.det_uniform((n, 27, 128, 128), 31, -1.0, 1.0))
    with torch.no_grad():
        wp, wf = ref(x); gp, gf = D(x.cuda())
    print("n",n,"prob err", (gp.cpu()-wp).abs().max().item())
    for i,(g,w) in enumerate(zip(gf,wf)):
        d=(g.cpu()-w); peak=w.abs().max().item()
        print(f" f{i+1}: max rel {d.abs().max().item()/peak:.4f} rms/peak {d.pow(2).mean().sqrt().item()/peak:.5f} psnr {20*math.log10(peak/d.pow(2).mean().sqrt().item()):.1f} dB  rms/rms {d.pow(2).mean().sqrt().item()/w.pow(2).mean().sqrt().item():.4f}")
