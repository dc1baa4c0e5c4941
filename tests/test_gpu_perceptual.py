"""The perceptual-loss STAND-IN (tecogan_b200.perceptual) is labelled NON-PARITY: the reference's VGG
branch cannot run (code/ops.py:153-166, code/train.py:30-45), so there is no reference output to match.  These tests hold
the stand-in to its OWN definition: the same random-init VGG19-to-conv4_4 network evaluated with torch convolutions on
bf16-rounded operands, the loss value, the gradient that reaches the generator output, and its integration in the
training step (off by default: vgg_scaling > 0 raises as in the reference)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import synth, train_oracle as TO

pytestmark = pytest.mark.gpu


def _torch_loss(P, gen, target):
    """the module's definition with torch ops: bf16-rounded conv operands, f32 accumulation."""
    def feats(x):
        img = ((x + 1) / 2) * 255.0 - P.mean
        out, k, cur = {}, 0, img
        names = {(1, 1): "conv2_2", (2, 3): "conv3_4", (3, 3): "conv4_4"}
        for gi, group in enumerate(((64, 64), (128, 128), (256,) * 4, (512,) * 4)):
            if gi:
                cur = F.max_pool2d(cur, 2)
            for li in range(len(group)):
                L = P.layers[k]
                k += 1
                cur = F.relu(F.conv2d(cur.to(torch.bfloat16).float(), L.weight.to(torch.bfloat16).float(), L.bias, padding=1))
                cur = cur.to(torch.bfloat16).float()                     # activations are stored in bf16
                if (gi, li) in names:
                    out[names[(gi, li)]] = cur
        return out
    fg, ft = feats(gen), feats(target.detach())
    total = 0
    for name in ("conv2_2", "conv3_4", "conv4_4"):
        a = fg[name] / torch.sqrt((fg[name] ** 2).sum(1, keepdim=True) + 1e-12)
        b = ft[name] / torch.sqrt((ft[name] ** 2).sum(1, keepdim=True) + 1e-12)
        total = total + (1.0 - (a * b).sum(1).mean())
    return total


def test_stand_in_matches_its_own_definition():
    from tecogan_b200 import perceptual
    P = perceptual.PerceptualStandIn("cuda", seed=19)
    gen = torch.from_numpy(synth.det_uniform((2, 3, 32, 48), 5, 0.05, 0.95)).cuda().requires_grad_(True)
    tgt = torch.from_numpy(synth.det_uniform((2, 3, 32, 48), 6, 0.05, 0.95)).cuda()
    loss, per_layer = P.loss(gen, tgt)
    loss.backward()
    g_ours = gen.grad.clone()
    gen.grad = None
    want = _torch_loss(P, gen, tgt)
    want.backward()
    assert len(per_layer) == 3 and all(0.0 <= float(l) <= 2.0 for l in per_layer)
    assert abs(float(loss) - float(want)) <= 2e-2 * max(float(want), 1e-3), (float(loss), float(want))
    a, b = g_ours.double().flatten(), gen.grad.double().flatten()
    cos = float(a @ b / (a.norm() * b.norm() + 1e-30))
    assert cos >= 0.98 and 0.9 <= float(a.norm() / b.norm()) <= 1.1, (cos, float(a.norm() / b.norm()))
    # identical inputs -> cosine similarity 1 everywhere -> zero loss
    same, _ = P.loss(tgt, tgt)
    assert abs(float(same)) <= 1e-3


def test_train_step_with_the_stand_in_enabled():
    from tecogan_b200 import models, perceptual, train as T
    args = TO.default_train_args(num_resblock=2, discrim_resblocks=1, discrim_channels=64, vgg_scaling=0.2)
    torch.manual_seed(4)
    G, D = models.generator(3, args).cuda(), models.discriminator(args).cuda()
    og = torch.optim.Adam(G.parameters(), 1e-4)
    od = torch.optim.Adam(D.parameters(), 1e-4)
    r_in = torch.from_numpy(synth.det_uniform((1, 10, 3, 32, 32), 71, 0.0, 1.0)).cuda()
    r_tg = torch.from_numpy(synth.det_uniform((1, 10, 3, 128, 128), 72, 0.0, 1.0)).cuda()
    assert not perceptual.ENABLED
    with pytest.raises(NotImplementedError):
        T.TecoGAN(r_in, r_tg, D, G, args, 0, 0.0, 0.0, og, od)
    perceptual.ENABLED = True
    try:
        w0 = G.output.weight.detach().clone()
        out = T.TecoGAN(r_in, r_tg, D, G, args, 0, 0.0, 0.0, og, od)
        names = out.update_list_name
        assert names[names.index("l2_warp_loss") + 1:names.index("t_adversarial_loss")] == ["vgg_loss_2", "vgg_loss_3", "vgg_loss_4", "vgg_all"]
        vals = {k: float(v) for k, v in zip(names, out.update_list)}
        assert all(v == v for v in vals.values()) and 0.0 < vals["vgg_all"] < 6.0
        assert (G.output.weight.detach() - w0).abs().max().item() > 0
        # the perceptual term reaches the generator: its gradient differs from the content-only gradient
        args2 = TO.default_train_args(num_resblock=2, discrim_resblocks=1, discrim_channels=64)
        torch.manual_seed(4)
        G2, D2 = models.generator(3, args2).cuda(), models.discriminator(args2).cuda()
        T.TecoGAN(r_in, r_tg, D2, G2, args2, 0, 0.0, 0.0, torch.optim.Adam(G2.parameters(), 1e-4), torch.optim.Adam(D2.parameters(), 1e-4))
        ga = torch.cat([p.grad.flatten() for p in G.parameters()])
        gb = torch.cat([p.grad.flatten() for p in G2.parameters()])
        assert (ga - gb).abs().max().item() > 1e-6 * gb.abs().max().item()
    finally:
        perceptual.ENABLED = False
