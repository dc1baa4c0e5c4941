"""Host-side mirror of the reference ``code/train.py``: ``TecoGAN`` / ``FRVSR_Train`` with the reference's signature,
return type (the ``Network`` namedtuple, code/train.py:357-360) and logged quantities, running on the B200 kernels.

What changes relative to the reference's step (same arithmetic):
  * the recurrent generator loop (code/train.py:86-114) is one autograd node: T sequential device-side frames forward
    (fused flow-upscale + warp + space-to-depth + concat kernel, then the persistent tcgen05 frame kernel), ONE batched
    backward over all B*T frames (the reference detaches every generator input, :90,108);
  * the discriminator's 27-channel inputs (:139-198: warp of 9-frame groups, centre crop + zero pad, LR resize, concat)
    come from one fused kernel per branch (tg_disc_input_assemble), the velocity field computed on the fly;
  * parameter gradients are accumulated by the wgrad kernels directly in one flat f32 bucket per network
    (tecogan_b200.parallel); under torch.distributed each bucket is all-reduced asynchronously right after its
    backward, overlapping the other network's backward;
  * no ``.cpu()`` hops (:294-301) — every scalar stays on the device.
Supported: everything the reference itself can run — its default flags and pingpang=True (code/train.py:56-62,153-156,
275-285).  Dt_mergeDs=False feeds a 9-channel tensor to the 27-channel discriminator and GAN_FLAG=False reads an unbound
t_adversarial_loss: both raise inside the reference (probed) and raise here.  vgg_scaling > 0 selects
the reference's VGG branch, which cannot run either; here it selects the labelled, non-parity stand-in of
tecogan_b200.perceptual when that module is enabled, else raises.
"""
import collections

import torch

import os
import warnings

from . import _native as _nt
from . import optim as _optim
from . import perceptual as _perceptual
from . import parallel as _par
from .models import *  # noqa: F401,F403  (reference: `from models import *`, code/train.py:1)

# Fused flat-bucket Adam (tecogan_b200.optim) for the stock torch.optim.Adam objects main.py:239-243 builds, and CUDA-graph
# capture of the whole step after GRAPH_WARMUP eager calls.  Both are on by default and fall back to the eager /
# stock-optimizer path when they do not apply; TG_TRAIN_FUSED_ADAM=0 / TG_TRAIN_GRAPH=0 switch them off (A/B measurements).
FUSED_ADAM = os.environ.get("TG_TRAIN_FUSED_ADAM", "1") != "0"
USE_CUDA_GRAPH = os.environ.get("TG_TRAIN_GRAPH", "1") != "0"
# The data-parallel step is captured as a SEQUENCE of graphs with the NCCL calls (two all-reduce launches, two waits) left
# eager between them (tecogan_b200.parallel.GraphSegments).  Capturing the collectives inside one graph also ran, but
# torch.distributed's process-group teardown then hung on a 2-GPU box.  TG_TRAIN_GRAPH_DP=0: eager step.
GRAPH_DATA_PARALLEL = os.environ.get("TG_TRAIN_GRAPH_DP", "1") != "0"
GRAPH_WARMUP = 2
PAIR_DISCRIMINATOR_PASSES = os.environ.get("TG_TRAIN_PAIR_D", "1") != "0"   # real + fake discriminator pass as one batch

VGG_MEAN = [123.68, 116.78, 103.94]          # code/train.py:6
identity = torch.nn.Identity()               # code/train.py:7

scaler = None                                # code/train.py:9 (created on first use: needs a CUDA device)


def _scaler():
    global scaler
    if scaler is None:
        scaler = torch.amp.GradScaler("cuda")
    return scaler


class EMA(torch.nn.Module):
    """code/train.py:13-26"""

    def __init__(self, mu):
        super().__init__()
        self.mu = mu
        self.shadow = {}

    def register(self, name, val):
        self.shadow[name] = val.clone()

    def forward(self, name, x):
        assert name in self.shadow
        new_average = self.mu * x + (1.0 - self.mu) * self.shadow[name]
        self.shadow[name] = new_average.clone()
        return new_average


def VGG19_slim(input, reuse, deep_list=None, norm_flag=True):
    """code/train.py:30-45 cannot run in the reference (VGG19() lacks its required arguments, torch.min(...) + float;
    SURVEY.md 8c) and vgg_scaling defaults to -0.002 (main.py:98), so the branch is never taken.  With the labelled,
    non-parity stand-in enabled (tecogan_b200.perceptual.ENABLED) TecoGAN uses that instead; otherwise this raises."""
    raise NotImplementedError("the reference's VGG perceptual branch is unrunnable (SURVEY.md 8c); keep vgg_scaling <= 0 or enable "
                              "the labelled non-parity stand-in (tecogan_b200.perceptual.ENABLED = True)")


Network = collections.namedtuple('Network', 'gen_output, learning_rate, update_list, update_list_name, update_list_avg, '
                                            'global_step, d_loss, gen_loss, fnet_loss ,tb, target')


def _check_args(args, r_inputs):
    if not getattr(args, "Dt_mergeDs", True):
        # code/train.py:183-184: discriminator_F(real_warp) with the 9-channel cropped tensor -> "expected input ... to
        # have 27 channels" in the reference's own conv (code/models.py:102); same exception type here
        raise RuntimeError("TecoGAN: Dt_mergeDs=False feeds a 9-channel input to the 27-channel discriminator "
                           "(code/train.py:183-184 raises in the reference as well)")
    if float(getattr(args, "vgg_scaling", -1.0)) > 0.0 and not _perceptual.ENABLED:
        VGG19_slim(None, None)
    if r_inputs.dim() != 5 or r_inputs.shape[2] != 3 or r_inputs.shape[3] != r_inputs.shape[4]:
        raise RuntimeError(f"TecoGAN: r_inputs must be [B,T,3,crop,crop], got {tuple(r_inputs.shape)}")
    if int(args.crop_size) != r_inputs.shape[3] or int(args.RNN_N) != r_inputs.shape[1]:
        raise RuntimeError("TecoGAN: args.crop_size / args.RNN_N do not match r_inputs")


def discriminator_inputs(r_inputs, r_targets, gen_tb, args):
    """code/train.py:130-198 -> (real_input, fake_input), each [t_batch,27,4c,4c] f32.
    r_inputs [B,T,3,c,c], r_targets [B,T,3,4c,4c], gen_tb = generator outputs FRAME-major [T,B,3,4c,4c]."""
    lib = _nt.lib()
    b, t, _, c, _ = r_inputs.shape
    hc = 4 * c
    ts = 3 * (t // 3)                                          # :130
    tb = b * ts // 3                                           # :135
    off = 0
    if args.crop_dt < 1.0:                                     # :160-164
        off = (hc - int(hc * args.crop_dt)) // 2
    lr9 = r_inputs[:, :ts].contiguous()                        # :176-177, [B,ts,3,c,c] == [tb,9,c,c]
    tgt9 = r_targets[:, :ts].contiguous()                      # :133-134,175
    # LR planes the reference up-scales into T_vel (:139-158); class = frame % 3
    gsrc = torch.zeros((b, ts // 3, 3, 2, c, c), dtype=torch.float32, device=r_inputs.device)
    gsrc[:, :, 0] = r_inputs[:, 0:ts:3, 0:2]                   # VPre  = gen_flow[:, 0:ts:3]           (:147,153)
    raw_next = 0
    if not getattr(args, "pingpang", False):
        back = torch.cat((r_inputs[:, 2:ts:3], r_inputs[:, 1:ts:3]), dim=1).reshape(tb, 6, c, c)     # :139-141
        gsrc[:, :, 2] = back[0:b].reshape(b, ts // 3, 2, c, c)     # VNxt = preprocess(up4(4 * back[0:B]))  (:143-145,149)
    else:
        # VNxt = flip(gen_flow, t)[:, 1:ts:3] (:155), gen_flow[:, i] = up4(4 * r_inputs[:, i])[:, 0:2] -> frames t-2-1, t-2-4, ...
        idx = torch.arange(t - 3, -1, -3, device=r_inputs.device)[: ts // 3]
        gsrc[:, :, 2] = r_inputs[:, idx, 0:2]
        raw_next = 2                                           # no preprocess() on this branch
    outs = []
    for src, sb, st_, fp16 in ((tgt9, ts * 3 * hc * hc, 3 * hc * hc, 0),            # real: grid_sample(t_targets, T_vel)   (:165)
                               (gen_tb.detach(), 3 * hc * hc, b * 3 * hc * hc, 1)):  # fake: grid_sample(t_gen, T_vel.half()) (:187)
        out = torch.empty((tb, 27, hc, hc), dtype=torch.float32, device=r_inputs.device)
        _nt.check(lib.tg_disc_input_assemble(_nt.ptr(tgt9), _nt.ptr(src), sb, st_, ts, _nt.ptr(gsrc), _nt.ptr(lr9), _nt.ptr(out),
                                             tb, c, c, off, fp16 | raw_next, _nt.stream_ptr()))
        outs.append(out)
    return outs[0], outs[1]


def _warp_loss(r_inputs):
    """code/train.py:71-85,247-249 — logged only: LR frame t-1 sampled with LR_t[:, 0:2] re-viewed as a grid (f32)."""
    b, t, _, c, _ = r_inputs.shape
    pre = r_inputs[:, :-1].reshape(b * (t - 1), 3, c, c)
    cur = r_inputs[:, 1:]
    grid = cur[:, :, 0:2].reshape(b * (t - 1), c, c, 2)
    s_warp = torch.nn.functional.grid_sample(pre, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return torch.mean(torch.sum(torch.square(cur.reshape(b * (t - 1), 3, c, c) - s_warp), dim=[3]))


def _dt_ratio_host(args, global_step):
    """code/train.py:291-292 on the host (all inputs are host scalars)."""
    return min(float(args.Dt_ratio_max), float(args.Dt_ratio_0) + float(args.Dt_ratio_add) * float(global_step))


def TecoGAN(r_inputs, r_targets, discriminator_F, generator_F, args, Global_step, counter1, counter2, optimizer_g,
            optimizer_d, GAN_FLAG=True, _dt_ratio_dev=None):
    """code/train.py:49-370.  `_dt_ratio_dev` (internal): a device scalar holding Dt_ratio for this step, so that a captured
    CUDA graph of this function follows Dt_ratio_add across replays."""
    if not GAN_FLAG:
        # code/train.py:293 reads t_adversarial_loss, which only the GAN_FLAG branch (:287-291) defines
        raise UnboundLocalError("TecoGAN: GAN_FLAG=False leaves t_adversarial_loss unbound (code/train.py:293 raises in "
                                "the reference as well; FRVSR_Train always passes True)")
    r_inputs = _nt.require_cuda_f32(r_inputs, "TecoGAN(r_inputs)")
    r_targets = _nt.require_cuda_f32(r_targets, "TecoGAN(r_targets)")
    _check_args(args, r_inputs)
    Global_step += 1                                                                       # :52
    rnn_n = r_inputs.shape[1]
    pingpang = bool(getattr(args, "pingpang", False))
    if pingpang:                                                                           # :56-62: forward + reversed clip
        r_inputs = torch.cat([r_inputs, torch.flip(r_inputs, dims=[1])[:, 1:]], dim=1)
        r_targets = torch.cat([r_targets, torch.flip(r_targets, dims=[1])[:, 1:]], dim=1)
    b, t, _, c, _ = r_inputs.shape
    hc = 4 * c
    learning_rate = args.learning_rate

    # flat gradient buckets first: the forward nodes below then hand their parameter gradients to the bucket instead of
    # autograd's per-parameter accumulators (tecogan_b200.models._grad_inputs)
    _par.bind_flat_grads(generator_F)
    _par.bind_flat_grads(discriminator_F)

    # ---- generator: all frames (:86-114)
    gen_tb = generator_F.forward_clip_train(r_inputs)                                      # [T,B,3,hc,hc]
    gen_outputs = gen_tb.transpose(0, 1)                                                   # [B,T,3,hc,hc] (view)
    update_list, update_list_name = [], []

    # ---- discriminator on real / fake triplets (:130-199)
    real_in, fake_in = discriminator_inputs(r_inputs, r_targets, gen_tb, args)
    if PAIR_DISCRIMINATOR_PASSES and hasattr(discriminator_F, "forward_pair"):
        # :181 and :199 (input detached) as one batch: same results as two calls (BatchNorm statistics per pass, running
        # statistics updated real first), half the launches
        (tdiscrim_real_output, real_layers), (tdiscrim_fake_output, fake_layers) = discriminator_F.forward_pair(real_in, fake_in)
    else:
        tdiscrim_real_output, real_layers = discriminator_F(real_in)                       # :181
        tdiscrim_fake_output, fake_layers = discriminator_F(fake_in)                       # :199 (input detached)

    # ---- layer losses (:203-232), detached on both sides
    sum_layer_loss = 0
    if args.D_LAYERLOSS:
        layer_norm = [12.0, 14.0, 24.0, 100.0]
        for i, (rl, fl) in enumerate(zip(real_layers, fake_layers)):
            layer_loss = torch.mean(torch.sum(torch.abs(rl.detach() - fl.detach()), dim=[3]))
            update_list.append(layer_loss)
            update_list_name.append("D_layer_%d_loss" % i)
            sum_layer_loss = sum_layer_loss + 0.02 * layer_loss / layer_norm[i]
        update_list.append(sum_layer_loss)
        update_list_name.append("D_layer_loss_sum")

    # ---- content / warp losses (:235-249)
    content_loss = torch.mean(torch.sum(torch.square(gen_outputs - r_targets), dim=[4]))   # == mean over [B*T,3,hc] rows (:239)
    update_list.append(content_loss)
    update_list_name.append("l2_content_loss")
    update_list.append(_warp_loss(r_inputs))
    update_list_name.append("l2_warp_loss")
    if float(getattr(args, "vgg_scaling", -1.0)) > 0.0:                                    # :124-127,253-273 -> the labelled stand-in
        bt = b * t
        vgg_loss, vgg_layers = _perceptual.get(r_inputs.device).loss(gen_outputs.reshape(bt, 3, hc, hc), r_targets.reshape(bt, 3, hc, hc))
        update_list += list(vgg_layers) + [vgg_loss]
        update_list_name += ["vgg_loss_%d" % (i + 2) for i in range(len(vgg_layers))] + ["vgg_all"]
    else:
        vgg_loss = None
    pploss = None
    if pingpang:                                                                           # :275-285
        pploss = torch.mean(torch.abs(gen_outputs[:, 0:rnn_n - 1] - torch.flip(gen_outputs, dims=[1])[:, :rnn_n - 1]))
        update_list.append(pploss)
        update_list_name.append("PingPang")

    # ---- adversarial terms (:288-301).  gen_loss, fnet_loss and content_loss are ONE tensor in the reference (:243-244),
    # updated in place: the adversarial term lands twice and the logged l2_content_loss equals All_loss_Gen.  Both extra
    # terms are detached: the generator's gradient is the content loss's alone.
    t_adversarial_loss = torch.mean(-torch.log(tdiscrim_fake_output.detach() + args.EPS))
    d_adversarial_loss = torch.mean(-torch.log(tdiscrim_fake_output + args.EPS))
    dt_ratio = torch.tensor(_dt_ratio_host(args, Global_step), dtype=torch.float32)          # :291-292 (a CPU tensor there too)
    dt_mul = _dt_ratio_dev if _dt_ratio_dev is not None else float(dt_ratio)
    gen_loss = content_loss
    fnet_loss = content_loss
    if vgg_loss is not None:                               # :267-268 (with gradient to the generator)
        gen_loss += args.vgg_scaling * vgg_loss
    if pploss is not None and args.pp_scaling > 0:         # :281-283, NOT detached: the generator also descends the ping-pong term
        gen_loss += pploss * args.pp_scaling
        fnet_loss += pploss * args.pp_scaling
    gen_loss += args.ratio * t_adversarial_loss
    fnet_loss += args.ratio * t_adversarial_loss
    update_list.append(t_adversarial_loss)
    update_list_name.append("t_adversarial_loss")
    if args.D_LAYERLOSS:
        gen_loss += sum_layer_loss * dt_mul

    # ---- discriminator loss (:303-322)
    t_discrim_fake_loss = torch.log(1 - tdiscrim_fake_output + args.EPS)
    t_discrim_real_loss = torch.log(tdiscrim_real_output + args.EPS)
    t_discrim_loss = torch.mean(-(t_discrim_fake_loss + t_discrim_real_loss))
    t_balance = torch.mean(t_discrim_real_loss) + d_adversarial_loss
    update_list += [t_discrim_loss, torch.mean(tdiscrim_real_output), torch.mean(tdiscrim_fake_output)]
    update_list_name += ["t_discrim_loss", "t_discrim_real_output", "t_discrim_fake_output"]
    discrim_loss = t_discrim_loss

    # ---- EMA bookkeeping (:324-332): one shadow chained through the list, started from zero -> a lower-triangular
    # weighting of the stacked scalars, done as one small mat-vec instead of 3 kernels per entry
    tb = 0.99 * t_balance.detach()
    update_list += [gen_loss]
    update_list_name += ["All_loss_Gen"]
    vals = torch.stack([v.detach().float().reshape(()) for v in update_list])
    k = vals.numel()
    idx = torch.arange(k, device=vals.device)
    expo = (idx[:, None] - idx[None, :]).clamp(min=0).float()
    tri = torch.where(idx[:, None] >= idx[None, :], 0.99 * torch.pow(torch.full((), 0.01, device=vals.device), expo),
                      torch.zeros((), device=vals.device))
    update_list_avg = list(tri @ vals)

    # ---- backward + optimizer steps (:335-342).  Reference order: G backward, G step, update, D backward, D step, update.
    # Here both backward passes run first (each followed by its asynchronous gradient all-reduce when data-parallel),
    # then the steps in the reference's order; should the first update() change the loss scale, the discriminator's
    # gradients are rescaled by the (power-of-two) ratio, which is what scaling its loss with the new scale produces.
    # The ratio stays on the device (GradScaler.get_scale() is a host synchronisation; scale(1) is not): with a fused
    # optimizer the whole step enqueues without the host ever waiting for the GPU.
    sc = _scaler()
    sync_g, sync_d = _par.GradSync(), _par.GradSync()
    fa_g = _optim.FlatAdam.adopt(generator_F, optimizer_g) if FUSED_ADAM else None
    fa_d = _optim.FlatAdam.adopt(discriminator_F, optimizer_d) if FUSED_ADAM else None
    optimizer_g.zero_grad()
    g_bucket = _par.zero_flat_grads(generator_F)
    one = torch.ones((), dtype=torch.float32, device=g_bucket.device)
    scale0 = sc.scale(one) if sc.is_enabled() else None
    sc.scale(gen_loss).backward()
    sync_g.start(g_bucket)
    optimizer_d.zero_grad()
    d_bucket = _par.zero_flat_grads(discriminator_F)
    sc.scale(discrim_loss).backward()
    sync_d.start(d_bucket)
    sync_g.finish()
    if fa_g is not None and fa_d is not None:
        # tecogan_b200.optim: inf/NaN check, Adam, GradScaler.update() and the bf16 re-pack as repo kernels on the flat
        # buckets; lr / step / loss scale / found_inf are device scalars (nothing here reads the GPU back).  Both gradients
        # were scaled by the loss scale BEFORE the first update(), so both are unscaled by 1 / that scale.
        if not torch.cuda.is_current_stream_capturing():
            fa_g.refresh_lr()
            fa_d.refresh_lr()
        inv0 = torch.reciprocal(scale0) if scale0 is not None else None
        skw = dict(scale=sc._scale, growth_tracker=sc._growth_tracker, growth=sc.get_growth_factor(), backoff=sc.get_backoff_factor(),
                   interval=sc.get_growth_interval()) if scale0 is not None else {}
        fa_g.step_kernels(g_bucket, inv0, **skw)
        sync_d.finish()
        fa_d.step_kernels(d_bucket, inv0, **skw)
    else:
        sc.step(optimizer_g)
        sc.update()
        sync_d.finish()
        if scale0 is not None:
            d_bucket.mul_(sc.scale(one) / scale0)              # exactly 1.0 unless update() just changed the scale
        sc.step(optimizer_d)
        sc.update()

    update_list_avg += [tb, dt_ratio]
    update_list_name += ["t_balance", "Dst_ratio"]
    update_list_avg += [counter1, counter2]
    update_list_name += ["withD_counter", "w_o_D_counter"]
    # callers .view() the generator output (main.py:287-289): hand it back contiguous in the reference's [B,T,...] order
    return Network(gen_output=gen_outputs.contiguous(), learning_rate=learning_rate, update_list=update_list,
                   update_list_name=update_list_name, update_list_avg=update_list_avg, global_step=Global_step,
                   d_loss=discrim_loss, gen_loss=gen_loss, fnet_loss=fnet_loss, tb=tb, target=real_in)


class _GraphedStep:
    """One captured CUDA graph of TecoGAN() for a fixed (networks, optimizers, batch shape, flags): the ~500 kernel
    launches of a step (cfg4) become one graph launch.  Inputs are copied into static buffers, Dt_ratio and the two
    learning rates are device scalars refreshed before every replay, everything else (weights, Adam moments, loss scale,
    BatchNorm running statistics, step counters) lives in device memory the graph updates in place.  The returned
    tensors are the graph's static outputs: valid until the next call (main.py consumes them immediately, :276-294)."""

    def __init__(self, r_inputs, r_targets):
        self.static_in = torch.empty_like(r_inputs)
        self.static_tg = torch.empty_like(r_targets)
        self.dt = torch.zeros((), dtype=torch.float32, device=r_inputs.device)
        self.graph = None
        self.out = None
        self.calls = 0
        self.failed = False
        self.kernels = 0                # repo-kernel launches recorded in the graph (tg_launch_count over the capture)

    def run(self, r_inputs, r_targets, args, D, G, step, c1, c2, og, od):
        self.static_in.copy_(r_inputs)
        self.static_tg.copy_(r_targets)
        self.dt.fill_(_dt_ratio_host(args, step + 1))
        for net, opt in ((G, og), (D, od)):
            _optim.FlatAdam.adopt(net, opt).refresh_lr()
        if self.graph is None:
            torch.cuda.synchronize()
            k0 = _nt.lib().tg_launch_count()
            if _par.world_size() > 1:
                # data-parallel: graph segments with the NCCL calls eager between them, captured on a side stream
                seg = _par.GraphSegments()
                side = torch.cuda.Stream(device=self.static_in.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    _par._segments = seg
                    seg.begin()
                    try:
                        self.out = TecoGAN(self.static_in, self.static_tg, D, G, args, step, c1, c2, og, od, _dt_ratio_dev=self.dt)
                    finally:
                        _par._segments = None
                    seg.end()
                torch.cuda.current_stream().wait_stream(side)
                self.graph = seg
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="global"):
                    self.out = TecoGAN(self.static_in, self.static_tg, D, G, args, step, c1, c2, og, od, _dt_ratio_dev=self.dt)
                self.graph = g
            self.kernels = int(_nt.lib().tg_launch_count() - k0)
        else:
            global replayed_launches
            replayed_launches += self.kernels   # (the capturing call's kernels were counted by tg_launch_count itself)
        self.graph.replay()
        out = self.out
        n = len(out.update_list)
        avg = list(out.update_list_avg[:n]) + [out.tb, torch.tensor(_dt_ratio_host(args, step + 1), dtype=torch.float32), c1, c2]
        return out._replace(global_step=step + 1, update_list_avg=avg, learning_rate=args.learning_rate)


_graphs = {}
replayed_launches = 0                   # repo kernels executed through graph replays (bench.py's gpu_launches adds this)


def _graph_key(r_inputs, r_targets, args, D, G, og, od):
    flags = tuple((k, getattr(args, k, None)) for k in ("pingpang", "crop_dt", "Dt_mergeDs", "D_LAYERLOSS", "vgg_scaling", "pp_scaling",
                                                         "ratio", "EPS", "RNN_N", "crop_size", "Dt_ratio_max", "Dt_ratio_0", "Dt_ratio_add"))
    return (id(G), id(D), id(og), id(od), tuple(r_inputs.shape), tuple(r_targets.shape), r_inputs.device, flags)


def FRVSR_Train(r_inputs, r_targets, args, discriminator_F, generator_F, step, counter1, counter2, optimizer_g,
                optimizer_d):
    """code/train.py:374-377.  After GRAPH_WARMUP eager calls with the same networks / optimizers / shapes the step is
    captured in a CUDA graph and replayed (single process, fused Adam adoptable, TG_TRAIN_GRAPH != 0)."""
    if (USE_CUDA_GRAPH and FUSED_ADAM and (_par.world_size() == 1 or GRAPH_DATA_PARALLEL) and isinstance(r_inputs, torch.Tensor) and r_inputs.is_cuda
            and _optim.FlatAdam.adoptable(generator_F, optimizer_g) and _optim.FlatAdam.adoptable(discriminator_F, optimizer_d)):
        key = _graph_key(r_inputs, r_targets, args, discriminator_F, generator_F, optimizer_g, optimizer_d)
        gs = _graphs.get(key)
        if gs is None:
            if len(_graphs) > 8:
                _graphs.clear()
            gs = _graphs[key] = _GraphedStep(_nt.require_cuda_f32(r_inputs, "FRVSR_Train(r_inputs)"),
                                             _nt.require_cuda_f32(r_targets, "FRVSR_Train(r_targets)"))
        gs.calls += 1
        if gs.calls > GRAPH_WARMUP and not gs.failed:
            try:
                return gs.run(r_inputs, r_targets, args, discriminator_F, generator_F, step, counter1, counter2, optimizer_g,
                              optimizer_d)
            except Exception as e:
                if gs.graph is not None:
                    raise                               # a captured graph that fails on replay is a real error
                # A capture that fails in capture_end leaves torch's capture bookkeeping (current stream, allocator pool, RNG
                # state) half-open: continuing eagerly in this process is not safe.  Fail loudly with the way out.
                gs.failed = True
                raise RuntimeError("tecogan_b200.train: CUDA-graph capture of the training step failed; set TG_TRAIN_GRAPH=0 "
                                   "(or tecogan_b200.train.USE_CUDA_GRAPH = False) to run the step eagerly") from e
    return TecoGAN(r_inputs, r_targets, discriminator_F, generator_F, args, step, counter1, counter2, optimizer_g,
                   optimizer_d)
