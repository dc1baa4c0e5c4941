"""torch-CPU fp32 restatement of the reference training step (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/code/train.py:49-370 (``TecoGAN``) for the configurations the reference can run: its default
flags (main.py:98-125: pingpang=False, crop_dt=0.75, Dt_mergeDs=True, D_LAYERLOSS=True, vgg_scaling<0) and
pingpang=True (train.py:56-62,153-156,275-285).  The VGG branch, Dt_mergeDs=False and GAN_FLAG=False crash in the
reference itself and are not restated.  kind = "port": pure Python over
PyTorch, restated over torch CPU ops.  PINNED by tests/golden/train.npz, train_cfg5.npz (64x64 crops, fc = 192 features) and train_pingpang.npz,
which oracle/make_golden.py writes by running the unmodified reference ``train.FRVSR_Train`` on CPU (``.cuda()`` patched to identity, grid cast to the image dtype).

The step is split into the same three stages the B200 implementation has, so tests can compare stage by stage:
``generator_loop`` (train.py:67-114), ``discriminator_inputs`` (train.py:130-198) and ``train_step`` (losses
train.py:203-329, the two backward passes and optimizer steps train.py:335-342).
"""
import types

import torch
import torch.nn.functional as F

from . import tecogan_oracle as O


def default_train_args(**kw):
    """argparse defaults read by the step (main.py:60-64,79,85,98-125)."""
    d = dict(num_resblock=16, discrim_resblocks=4, discrim_channels=128, RNN_N=10, crop_size=32, pingpang=False,
             learning_rate=1e-4, vgg_scaling=-0.002, crop_dt=0.75, Dt_mergeDs=True, D_LAYERLOSS=True, EPS=1e-12,
             ratio=0.01, Dt_ratio_max=1.0, Dt_ratio_0=1.0, Dt_ratio_add=0.0, pp_scaling=1.0, beta=0.9, adameps=1e-8)
    d.update(kw)
    return types.SimpleNamespace(**d)


def flow_from_lr(lr_frames):
    """train.py:71-77: the 'flow' is the first two channels of upscale_four(4 * LR).  lr_frames [M,3,h,w] -> [M,2,4h,4w]."""
    return O.upscale_four(lr_frames * 4.0)[:, 0:2]


def generator_loop(G, r_inputs):
    """train.py:86-114.  r_inputs [B,T,3,c,c] -> gen_outputs [B,T,3,4c,4c]; every generator input is detached
    (train.py:90,108), so no gradient flows from frame t to frame t-1."""
    b, t, _, c, _ = r_inputs.shape
    flow = flow_from_lr(r_inputs[:, :-1].reshape(b * (t - 1), 3, c, c)).reshape(b, t - 1, 2, 4 * c, 4 * c)   # :71-77
    x0 = torch.cat((r_inputs[:, 0], torch.zeros(b, 48, c, c)), dim=1)                                         # :86-88
    prev = G(x0.detach()).view(b, 3, 4 * c, 4 * c)                                                           # :90-91
    outs = [prev]
    for i in range(t - 1):                                                                                   # :94
        grid = flow[:, i].reshape(b, 4 * c, 4 * c, 2)         # a raw re-view of [2,Ho,Wo] memory, as the reference (:96)
        warped = O.warp(prev, grid)                                                                          # :98
        x = torch.cat((r_inputs[:, i + 1], O.space_to_depth(O.deprocess(warped), 4)), dim=1)                 # :101-107
        prev = G(x.detach()).view(b, 3, 4 * c, 4 * c)                                                        # :108-111
        outs.append(prev)
    return torch.stack(outs, dim=1), flow                                                                    # :113-114


def _crop_pad(x, off):
    """train.py:160-174: resized_crop to the centre window (same size: no resampling) then zero pad back."""
    if off == 0:
        return x
    return F.pad(x[..., off:-off, off:-off], (off, off, off, off))


def discriminator_inputs(r_inputs, r_targets, gen_outputs, flow, args):
    """train.py:130-198 for Dt_mergeDs=True (both pingpang settings).  Returns (real_input, fake_input) [t_batch,27,4c,4c]."""
    b, t, _, c, _ = r_inputs.shape
    hc = 4 * c
    ts = 3 * (t // 3)                                                                                        # :130
    tb = b * ts // 3                                                                                         # :135
    t_gen = gen_outputs[:, :ts].reshape(b * ts, 3, hc, hc)                                                   # :131-132
    t_tgt = r_targets[:, :ts].reshape(b * ts, 3, hc, hc)                                                     # :133-134
    v_pre = flow[:, 0:ts:3]                                                                                  # :147,153
    if not getattr(args, "pingpang", False):
        back = torch.cat((r_inputs[:, 2:ts:3], r_inputs[:, 1:ts:3]), dim=1).reshape(tb, 6, c, c)             # :139-141
        flow_back = O.upscale_four(back[0:b] * 4.0).reshape(b, ts // 3, 2, hc, hc)                           # :143-145
        v_nxt = O.preprocess(flow_back)                                                                      # :149
    else:
        v_nxt = torch.flip(flow, dims=[1])[:, 1:ts:3]                                                        # :155 (no preprocess)
    t_vel = torch.stack([v_pre, torch.zeros_like(v_pre), v_nxt], dim=2).reshape(b * ts, hc, hc, 2).detach()  # :156-158
    off = 0
    if args.crop_dt < 1.0:                                                                                   # :160-164
        off = (hc - int(hc * args.crop_dt)) // 2
    real_warp = F.grid_sample(t_tgt, t_vel, align_corners=False).reshape(tb, 9, hc, hc)                      # :165-167 (grid NOT fp16)
    real_warp = _crop_pad(real_warp, off)
    before = t_tgt.reshape(tb, 9, hc, hc)                                                                    # :175
    t_in = r_inputs[:, :ts].reshape(tb, 9, c, c)                                                             # :176-177
    input_hi = F.interpolate(t_in, size=(hc, hc), mode="bilinear", align_corners=False)                      # :178 (upsampling: antialias is a no-op)
    real_in = torch.cat((before, real_warp, input_hi), dim=1)                                                # :179
    fake_warp = O.warp(t_gen, t_vel).reshape(tb, 9, hc, hc)                                                  # :187-189 (grid .half())
    fake_warp = _crop_pad(fake_warp, off)
    fake_in = torch.cat((before, fake_warp, input_hi), dim=1)                                                # :197-198 (targets' before_warp again)
    return real_in, fake_in


def train_step(G, D, opt_g, opt_d, r_inputs, r_targets, args, global_step=0):
    """One TecoGAN step; returns a dict with the logged scalars (same names as update_list_name), the generator output,
    the discriminator's real input and the two losses.  Parameter gradients are left in ``.grad`` (unscaled: GradScaler
    is disabled on CPU), parameters are updated by the two Adam steps."""
    global_step += 1
    rnn_n = r_inputs.shape[1]
    if getattr(args, "pingpang", False):                                                                     # :56-62
        r_inputs = torch.cat([r_inputs, torch.flip(r_inputs, dims=[1])[:, 1:]], dim=1)
        r_targets = torch.cat([r_targets, torch.flip(r_targets, dims=[1])[:, 1:]], dim=1)
    b, t, _, c, _ = r_inputs.shape
    gen_outputs, flow = generator_loop(G, r_inputs)
    s_gen = gen_outputs.reshape(b * t, 3, 4 * c, 4 * c)                                                      # :116-118
    s_tgt = r_targets.reshape(b * t, 3, 4 * c, 4 * c)
    log = {}
    real_in, fake_in = discriminator_inputs(r_inputs, r_targets, gen_outputs, flow, args)
    p_real, real_layers = D(real_in)                                                                         # :181
    p_fake, fake_layers = D(fake_in.detach())                                                                # :199
    norms = [12.0, 14.0, 24.0, 100.0]                                                                        # :212
    sum_layer = 0
    for i, (rl, fl) in enumerate(zip(real_layers, fake_layers)):                                             # :213-224
        ll = torch.mean(torch.sum(torch.abs(rl.detach() - fl.detach()), dim=[3]))
        log["D_layer_%d_loss" % i] = ll
        sum_layer = sum_layer + 0.02 * ll / norms[i]
    log["D_layer_loss_sum"] = sum_layer
    content = torch.mean(torch.sum(torch.square(s_gen - s_tgt), dim=[3]))                                    # :239-241
    log["l2_content_loss"] = content
    # train.py:71-85,247-249: the logged warp loss (LR frame t vs frame t-1 sampled with LR_t[:, :2] re-viewed as a grid)
    pre = r_inputs[:, :-1].reshape(b * (t - 1), 3, c, c)
    cur = r_inputs[:, 1:]
    s_warp = F.grid_sample(pre, cur[:, :, 0:2].reshape(b * (t - 1), c, c, 2), align_corners=False)
    log["l2_warp_loss"] = torch.mean(torch.sum(torch.square(cur.reshape(b * (t - 1), 3, c, c) - s_warp), dim=[3]))
    pploss = None
    if getattr(args, "pingpang", False):                                                                     # :275-285
        first = gen_outputs[:, 0:rnn_n - 1]
        last_rev = torch.flip(gen_outputs, dims=[1])[:, :rnn_n - 1]
        pploss = torch.mean(torch.abs(first - last_rev))
        log["PingPang"] = pploss
    t_adv = torch.mean(-torch.log(p_fake.detach() + args.EPS))                                               # :289
    d_adv = torch.mean(-torch.log(p_fake + args.EPS))                                                        # :290
    dt_ratio = min(args.Dt_ratio_max, args.Dt_ratio_0 + args.Dt_ratio_add * float(global_step))              # :291-292
    # train.py:243-244,294-301: gen_loss, fnet_loss and content_loss are ONE tensor updated in place, so the adversarial
    # term is added twice and the logged l2_content_loss ends up equal to All_loss_Gen.  Values only: both extra terms are
    # detached, the generator's gradient is that of the content loss alone.
    gen_loss = content
    if pploss is not None and args.pp_scaling > 0:                                                           # :281-283 (NOT detached; added
        gen_loss += pploss * args.pp_scaling                                                                 #  twice: gen_loss is fnet_loss)
        gen_loss += pploss * args.pp_scaling
    gen_loss += args.ratio * t_adv
    gen_loss += args.ratio * t_adv
    log["t_adversarial_loss"] = t_adv
    gen_loss += sum_layer * dt_ratio
    fake_l = torch.log(1 - p_fake + args.EPS)                                                                # :305-308
    real_l = torch.log(p_real + args.EPS)
    d_loss = torch.mean(-(fake_l + real_l))
    t_balance = torch.mean(real_l) + d_adv                                                                   # :309
    log["t_discrim_loss"] = d_loss
    log["t_discrim_real_output"] = torch.mean(p_real)
    log["t_discrim_fake_output"] = torch.mean(p_fake)
    log["All_loss_Gen"] = gen_loss
    # train.py:324-332: EMA(0.99) started from zero, and ONE shadow chained through the whole list of scalars
    tb = 0.99 * t_balance
    avg, shadow = [], torch.zeros(())
    for v in log.values():
        shadow = 0.99 * v.detach() + 0.01 * shadow
        avg.append(shadow)
    opt_g.zero_grad()                                                                                        # :335-338
    gen_loss.backward()
    opt_g.step()
    opt_d.zero_grad()                                                                                        # :339-342
    d_loss.backward()
    opt_d.step()
    return dict(log={k: float(v.detach()) for k, v in log.items()}, log_avg=[float(v) for v in avg], tb=float(tb.detach()),
                dt_ratio=float(dt_ratio), gen_output=gen_outputs.detach(), target=real_in.detach(),
                d_loss=float(d_loss.detach()), gen_loss=float(gen_loss.detach()))
