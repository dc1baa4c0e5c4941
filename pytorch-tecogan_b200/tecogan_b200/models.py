"""Host-side mirror of the reference ``code/models.py`` (generator / discriminator interfaces).

``generator(gen_output_channels, args)`` keeps the reference's constructor, attribute tree
(``conv``, ``resids``, ``conv_trans``, ``output`` — hence identical ``state_dict`` keys,
shapes and default initialisation; code/models.py:61-76) and ``forward`` contract
(x [N,51,H,W] -> [N,3,4H,4W], NCHW-contiguous, values in (0,1); code/models.py:78-86), but
``forward`` runs the 41 convolutions as hand-written tcgen05 kernels through the C ABI.
The nn.Conv2d children are parameter containers only; they are never called.
"""
import torch
import torch.nn as nn

from . import _native as _nt
from . import parallel as _par
from .ops import *  # noqa: F401,F403  (reference star-import chain: code/models.py:1 -> dataloader -> ops)
from .ops import conv2, conv2_tran, lrelu, batchnorm, denselayer


def _grad_inputs(module, params):
    """Autograd inputs that make a B200 forward node differentiable w.r.t. the module's parameters.

    Plain use: the parameters themselves; the backward returns one gradient per parameter.
    Flat-bucket use (tecogan_b200.parallel.bind_flat_grads: every p.grad IS a view of the module's flat gradient bucket, the
    state tecogan_b200.train.TecoGAN sets up before its forward passes): ONE fresh scalar leaf.  The backward kernels add the
    parameter gradients straight into the bucket and return nothing, so autograd never touches the parameters' cached
    AccumulateGrad nodes - those remember the stream they were created on, and a node left over from an eager step on the
    default stream would make the engine synchronise the default stream with a capturing stream and invalidate a CUDA-graph
    capture of the step."""
    if _par.bucket_bound(module, params):
        return True, (torch.zeros((), dtype=torch.float32, device=params[0].device, requires_grad=True),)
    return False, tuple(params)


def residual_block(inputs, output_channel=64, stride=1):
    """code/models.py:54-58"""
    return nn.Sequential(conv2(inputs, 3, output_channel, stride, use_bias=True), nn.ReLU(),
                         conv2(output_channel, 3, output_channel, stride, use_bias=False))


class _GeneratorFn(torch.autograd.Function):
    """generator.forward with gradients (code/train.py:336 back-propagates the content loss through it): the forward
    keeps all activations in a per-call workspace, the backward runs the tcgen05 dgrad / wgrad kernels
    (tg_gen_backward).  The input is detached in the reference (code/train.py:90,108): no input gradient."""

    @staticmethod
    def forward(ctx, x, module, use_bucket, *params):
        lib = _nt.lib()
        x = x.detach().float().contiguous()
        n, _, h, w = x.shape
        nres = int(module.num)
        packed = module.packed_weights()
        ws = torch.empty(lib.tg_gen_train_workspace_bytes(n, h, w, nres), dtype=torch.uint8, device=x.device)
        out = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=x.device)
        _nt.check(lib.tg_gen_forward_train(_nt.ptr(packed), nres, _nt.ptr(x), _nt.ptr(out), _nt.ptr(ws), ws.numel(), n, h, w,
                                           _nt.stream_ptr()))
        ctx.module, ctx.ws, ctx.shape, ctx.use_bucket = module, ws, (n, h, w), use_bucket
        ctx.packed_dgrad = module.packed_dgrad_weights()
        ctx.save_for_backward(out)
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _nt.lib()
        (out,) = ctx.saved_tensors
        n, h, w = ctx.shape
        module = ctx.module
        nres = int(module.num)
        g = grad_out.float().contiguous()
        return (None, None, None) + _gen_backward(module, ctx.packed_dgrad, g, out, ctx.ws, n, h, w, ctx.use_bucket)


def _gen_backward(module, packed_dgrad, dout, out, ws, n, h, w, use_bucket):
    """tg_gen_backward over n images; returns the per-parameter gradients, or a single None (for the anchor leaf of
    _grad_inputs) when the forward ran in flat-bucket mode: p.grad are views of the bucket, the kernels add in place."""
    lib = _nt.lib()
    nres = int(module.num)
    params = module._param_list()
    if use_bucket and not _par.bucket_bound(module, params):
        raise RuntimeError("generator backward: the flat gradient bucket was unbound between forward and backward "
                           "(p.grad re-assigned?); re-run the forward")
    bucket = module._grad_bucket if use_bucket else None
    flat = bucket if bucket is not None else torch.zeros(lib.tg_gen_param_count(nres), dtype=torch.float32, device=out.device)
    _nt.check(lib.tg_gen_backward(_nt.ptr(packed_dgrad), nres, _nt.ptr(dout), _nt.ptr(out), _nt.ptr(flat), _nt.ptr(ws),
                                  ws.numel(), n, h, w, _nt.stream_ptr()))
    if bucket is not None:
        return (None,)
    grads, o = [], 0
    for p in params:
        grads.append(flat[o:o + p.numel()].view_as(p).to(p.dtype))
        o += p.numel()
    return tuple(grads)


class _GeneratorClipFn(torch.autograd.Function):
    """The recurrent generator loop of a training step (code/train.py:86-111) as ONE autograd node: the forward runs the
    T frames sequentially on the device (fused frame-input kernel + persistent frame kernel per frame, activations kept),
    the backward is a single batched pass over all B*T frames — the reference detaches every generator input
    (code/train.py:90,108), so the frames are independent in the backward direction."""

    @staticmethod
    def forward(ctx, lr, module, use_bucket, *params):
        lib = _nt.lib()
        lr = _nt.require_cuda_f32(lr.detach(), "forward_clip_train(r_inputs)")
        b, t, c, h, w = lr.shape
        nres = int(module.num)
        packed = module.packed_weights()
        ws = torch.empty(lib.tg_gen_train_workspace_bytes(b * t, h, w, nres), dtype=torch.uint8, device=lr.device)
        out = torch.empty((t, b, 3, 4 * h, 4 * w), dtype=torch.float32, device=lr.device)
        _nt.check(lib.tg_gen_clip_forward_train(_nt.ptr(packed), nres, _nt.ptr(lr), _nt.ptr(out), _nt.ptr(ws), ws.numel(),
                                                b, t, h, w, _nt.stream_ptr()))
        ctx.module, ctx.ws, ctx.shape, ctx.use_bucket = module, ws, (b * t, h, w), use_bucket
        ctx.packed_dgrad = module.packed_dgrad_weights()
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (out,) = ctx.saved_tensors
        n, h, w = ctx.shape
        g = grad_out.float().contiguous()
        grads = _gen_backward(ctx.module, ctx.packed_dgrad, g, out, ctx.ws, n, h, w, ctx.use_bucket)
        ctx.ws = None
        return (None, None, None) + grads


class generator(nn.Module):
    """code/models.py:61-86, B200-native forward and backward."""

    def __init__(self, gen_output_channels, args=None):
        super().__init__()
        if args is None:
            raise ValueError("No args is provided for generator")       # code/models.py:65-66
        if int(gen_output_channels) != 3:
            raise ValueError("tecogan_b200 generator supports gen_output_channels == 3 (the reference's only use)")
        self.conv = nn.Sequential(conv2(51, 3, 64, 1), nn.ReLU())
        self.num = args.num_resblock
        self.resids = nn.ModuleList([residual_block(64, 64, 1) for _ in range(int(self.num))])
        self.conv_trans = nn.Sequential(conv2_tran(64, 3, 64, stride=2, output_padding=1), nn.ReLU(),
                                        residual_block(64, 64, 1), residual_block(64, 128, 1),
                                        conv2_tran(128, 3, 128, stride=2, output_padding=1), nn.ReLU(),
                                        conv2(128, 3, 64, 1), nn.ReLU())
        self.output = conv2(64, 3, gen_output_channels, 1)
        self.amode = _nt.AMODE_FRAME
        self._packed = None
        self._packed_key = None
        self._ws = None

    # ------------------------------------------------------------------ packed-weight cache
    def _param_list(self):
        # state_dict iteration order == the flat layout tg_gen_pack expects
        return [p for _, p in self.named_parameters()]

    def invalidate_packed(self):
        """Drop the packed bf16 weight caches.  They are keyed on (data_ptr, tensor version); a write through ``p.data``
        bumps neither, so code that writes that way (EMA weight swaps, ``p.data.clamp_()``) calls this afterwards.
        ``load_state_dict`` and ``tecogan_b200.parallel.broadcast_parameters`` do it themselves."""
        self._packed_key = None
        self._packed_dgrad_key = None

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate_packed()
        return r

    def _flat_source(self, params):
        """the flat f32 parameter vector the pack kernels read: the bound flat buffer (tecogan_b200.optim.bind_flat_params:
        the parameters ARE views of it, no copy) or a gathered copy."""
        flat = getattr(self, "_flat_params", None)
        if flat is not None and flat.data_ptr() == params[0].data_ptr() and flat.device == params[0].device:
            return flat
        return torch.cat([p.detach().reshape(-1).float() for p in params])

    def repack_from_flat(self):
        """Re-derive the packed bf16 weights after tecogan_b200.optim.FlatAdam updated the flat parameter buffer in place
        (raw-pointer writes bump no tensor version: the caches keep their keys, their CONTENT is refreshed here).  Two
        batched launches, graph-capturable, no gather copy."""
        lib = _nt.lib()
        params = self._param_list()
        flat = self._flat_source(params)
        nres = int(self.num)
        st = _nt.stream_ptr(flat.device)
        if self._packed is not None and self._packed_key is not None:
            _nt.check(lib.tg_gen_pack(_nt.ptr(flat), nres, _nt.ptr(self._packed), st))
        if getattr(self, "_packed_dgrad", None) is not None and getattr(self, "_packed_dgrad_key", None) is not None:
            _nt.check(lib.tg_gen_pack_dgrad(_nt.ptr(flat), nres, _nt.ptr(self._packed_dgrad), st))

    def packed_weights(self):
        """bf16 MMA-ordered weight blocks + f32 biases; a derived cache, rebuilt whenever a
        parameter changed (optimizer step, load_state_dict) — tracked by tensor versions."""
        params = self._param_list()
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or key != self._packed_key:
            lib = _nt.lib()
            dev = params[0].device
            if dev.type != "cuda":
                raise RuntimeError("generator parameters must live on a CUDA device (call .cuda()); no CPU fallback")
            nres = int(self.num)
            flat = self._flat_source(params)
            assert flat.numel() == lib.tg_gen_param_count(nres), "parameter layout mismatch"
            if self._packed is None or self._packed.device != dev:
                self._packed = torch.empty(lib.tg_gen_packed_bytes(nres), dtype=torch.uint8, device=dev)
            _nt.check(lib.tg_gen_pack(_nt.ptr(flat), nres, _nt.ptr(self._packed), _nt.stream_ptr()))
            self._packed_key = key
        return self._packed

    def packed_dgrad_weights(self):
        """packed weights of the data-gradient convolutions (training only); same cache discipline."""
        params = self._param_list()
        key = tuple((p.data_ptr(), p._version) for p in params)
        if getattr(self, "_packed_dgrad", None) is None or key != getattr(self, "_packed_dgrad_key", None):
            lib = _nt.lib()
            nres = int(self.num)
            flat = self._flat_source(params)
            buf = getattr(self, "_packed_dgrad", None)
            if buf is None or buf.device != flat.device:
                buf = torch.empty(lib.tg_gen_packed_dgrad_bytes(nres), dtype=torch.uint8, device=flat.device)
            _nt.check(lib.tg_gen_pack_dgrad(_nt.ptr(flat), nres, _nt.ptr(buf), _nt.stream_ptr()))
            self._packed_dgrad, self._packed_dgrad_key = buf, key
        return self._packed_dgrad

    def _workspace(self, n, h, w, dev):
        need = _nt.lib().tg_gen_workspace_bytes(n, h, w)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._ws

    # ----------------------------------------------------------------------------- forward
    def forward(self, x, return_logits=False):
        if not x.is_cuda:
            raise RuntimeError("generator.forward: input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 51:
            raise RuntimeError(f"generator.forward: expected [N,51,H,W], got {tuple(x.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if x.requires_grad:
                raise NotImplementedError("tecogan_b200 generator: no gradient w.r.t. the input (the reference detaches it, "
                                          "code/train.py:90,108)")
            if return_logits:
                raise RuntimeError("return_logits is an inference-only probe")
            use_bucket, leaves = _grad_inputs(self, self._param_list())
            return _GeneratorFn.apply(x, self, use_bucket, *leaves)
        lib = _nt.lib()
        x = x.float().contiguous()
        n, _, h, w = x.shape
        packed = self.packed_weights()
        ws = self._workspace(n, h, w, x.device)
        x_nhwc = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=x.device)
        _nt.check(lib.tg_pack_nchw_to_nhwc64(_nt.ptr(x), _nt.ptr(x_nhwc), n, 51, h, w, _nt.stream_ptr()))
        out = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=x.device)
        logits = torch.empty_like(out) if return_logits else None
        _nt.check(lib.tg_gen_forward(_nt.ptr(packed), int(self.num), _nt.ptr(x_nhwc), _nt.ptr(out), _nt.ptr(logits),
                                     _nt.ptr(ws), ws.numel(), n, h, w, int(self.amode), _nt.stream_ptr()))
        return (out, logits) if return_logits else out

    def forward_clip_train(self, r_inputs):
        """code/train.py:86-114 on the device: r_inputs [B,T,3,H,W] -> generator outputs [T,B,3,4H,4W] (FRAME-major, f32,
        differentiable w.r.t. the parameters; ``.transpose(0, 1)`` gives the reference's [B,T,...] order)."""
        if r_inputs.dim() != 5 or r_inputs.shape[2] != 3:
            raise RuntimeError(f"forward_clip_train: expected [B,T,3,H,W], got {tuple(r_inputs.shape)}")
        use_bucket, leaves = _grad_inputs(self, self._param_list())
        return _GeneratorClipFn.apply(r_inputs, self, use_bucket, *leaves)

    @torch.no_grad()
    def infer_clip(self, r_inputs):
        """The whole recurrent loop of main.py:173-219 on the device:
        r_inputs [B,T,3,H,W] (CUDA, f32) -> [B,T,3,4H,4W] f32."""
        lib = _nt.lib()
        r = _nt.require_cuda_f32(r_inputs, "infer_clip(r_inputs)")
        b, t, c, h, w = r.shape
        if c != 3:
            raise RuntimeError("infer_clip: LR frames must have 3 channels")
        packed = self.packed_weights()
        ws = self._workspace(b, h, w, r.device)
        out = torch.empty((b, t, 3, 4 * h, 4 * w), dtype=torch.float32, device=r.device)
        _nt.check(lib.tg_gen_clip_forward(_nt.ptr(packed), int(self.num), _nt.ptr(r), _nt.ptr(out), _nt.ptr(ws),
                                          ws.numel(), b, t, h, w, int(self.amode), _nt.stream_ptr()))
        return out


def discriminator_block(inputs, output_channel, kernel_size, stride):
    """code/models.py:90-94"""
    return nn.Sequential(conv2(inputs, kernel_size, output_channel, stride, use_bias=False),
                         batchnorm(output_channel, is_training=True), lrelu(0.2))


class discriminator(nn.Module):
    """code/models.py:97-146, B200-native forward: tcgen05 3x3 / 4x4-stride-2 convolutions, BatchNorm (batch
    statistics, always train mode in the reference) + LeakyReLU / skip passes and the fc + sigmoid head run as
    hand-written kernels through the C ABI (tg_disc_forward).  The nn.Module children are parameter / buffer
    containers (identical state_dict keys, shapes and default initialisation); they are never called."""

    def __init__(self, args=None):
        super().__init__()
        if args is None:
            raise ValueError("No args is provided for discriminator")   # code/models.py:100-101
        ch = int(args.discrim_channels)
        nb = int(args.discrim_resblocks)
        self.conv = nn.Sequential(conv2(27, 3, 64, 1), lrelu(0.2))
        self.block1 = discriminator_block(64, 64, 4, 2)
        self.resids1 = nn.ModuleList([nn.Sequential(residual_block(64, 64, 1), batchnorm(64, True)) for _ in range(nb)])
        self.block2 = discriminator_block(64, ch, 4, 2)
        self.resids2 = nn.ModuleList([nn.Sequential(residual_block(ch, ch, 1), batchnorm(ch, True)) for _ in range(nb)])
        self.block3 = discriminator_block(ch, ch, 4, 2)
        self.resids3 = nn.ModuleList([nn.Sequential(residual_block(ch, ch, 1), batchnorm(ch, True)) for _ in range(nb)])
        self.block4 = discriminator_block(ch, 64, 4, 2)
        self.block5 = discriminator_block(64, 3, 4, 2)
        # 48 = 3*4*4 features of a 128x128 input (32x32 LR crops, code/models.py:123); larger crops
        # need 48*(crop/32)^2 (colab/README.md:15-22) — derived here, default unchanged.
        crop = int(getattr(args, "crop_size", 32) or 32)
        self.fc = denselayer(48 * max(1, crop // 32) ** 2, 1)
        self.nb, self.ch = nb, ch
        self._packed = None
        self._flat = None
        self._key = None
        self._ws = None

    def invalidate_packed(self):
        """Drop the packed bf16 weight caches (see generator.invalidate_packed)."""
        self._key = None
        self._packed_dgrad_key = None

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate_packed()
        return r

    def _bn_modules(self):
        mods = [self.block1[1]] + [r[1] for r in self.resids1] + [self.block2[1]] + [r[1] for r in self.resids2]
        mods += [self.block3[1]] + [r[1] for r in self.resids3] + [self.block4[1], self.block5[1]]
        return mods

    def _weights(self):
        """(flat f32 parameters in named_parameters order, packed bf16 conv blocks) — a derived cache rebuilt
        whenever a parameter changed (tensor versions)."""
        params = [p for _, p in self.named_parameters()]
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or key != self._key:
            lib = _nt.lib()
            dev = params[0].device
            if dev.type != "cuda":
                raise RuntimeError("discriminator parameters must live on a CUDA device (call .cuda()); no CPU fallback")
            bound = getattr(self, "_flat_params", None)
            if bound is not None and bound.data_ptr() == params[0].data_ptr() and bound.device == dev:
                flat = bound                        # the parameters ARE views of the flat buffer (tecogan_b200.optim)
            else:
                flat = torch.cat([p.detach().reshape(-1).float() for p in params])
            assert flat.numel() == lib.tg_disc_param_count(self.nb, self.ch, self.fc.in_features), "parameter layout mismatch"
            packed = self._packed
            if packed is None or packed.device != dev:
                packed = torch.empty(lib.tg_disc_packed_bytes(self.nb, self.ch), dtype=torch.uint8, device=dev)
            _nt.check(lib.tg_disc_pack(_nt.ptr(flat), self.nb, self.ch, _nt.ptr(packed), _nt.stream_ptr()))
            self._flat, self._packed, self._key = flat, packed, key
        return self._flat, self._packed

    def repack_from_flat(self):
        """see generator.repack_from_flat: refresh the packed conv blocks after the flat parameters were updated in place."""
        lib = _nt.lib()
        if self._packed is None or self._key is None or self._flat is None:
            return
        st = _nt.stream_ptr(self._flat.device)
        _nt.check(lib.tg_disc_pack(_nt.ptr(self._flat), self.nb, self.ch, _nt.ptr(self._packed), st))
        if getattr(self, "_packed_dgrad", None) is not None and getattr(self, "_packed_dgrad_key", None) is not None:
            _nt.check(lib.tg_disc_pack_dgrad(_nt.ptr(self._flat), self.nb, self.ch, _nt.ptr(self._packed_dgrad), st))

    def _dgrad_weights(self):
        """packed weights of the data-gradient convolutions (training only); rebuilt with the forward cache."""
        flat, _ = self._weights()
        if getattr(self, "_packed_dgrad_key", None) is None or self._packed_dgrad_key != self._key:
            lib = _nt.lib()
            buf = getattr(self, "_packed_dgrad", None)
            if buf is None or buf.device != flat.device:
                buf = torch.empty(lib.tg_disc_packed_dgrad_bytes(self.nb, self.ch), dtype=torch.uint8, device=flat.device)
            _nt.check(lib.tg_disc_pack_dgrad(_nt.ptr(flat), self.nb, self.ch, _nt.ptr(buf), _nt.stream_ptr()))
            self._packed_dgrad, self._packed_dgrad_key = buf, self._key
        return self._packed_dgrad

    def _run(self, x, ws, groups=1):
        """one tg_disc_forward launch sequence into workspace `ws`; returns (prob, feats).  groups = 2: x holds two
        independent passes back to back (BatchNorm statistics per pass)."""
        import ctypes
        lib = _nt.lib()
        n, _, h, w = x.shape
        flat, packed = self._weights()
        prob = torch.empty((n, 1), dtype=torch.float32, device=x.device)
        shapes = [(n, 64, h // 2, w // 2), (n, self.ch, h // 4, w // 4), (n, self.ch, h // 8, w // 8), (n, 64, h // 16, w // 16)]
        feats = [torch.empty(s, dtype=torch.float32, device=x.device) for s in shapes]
        feat_ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in feats])
        bns = self._bn_modules()
        run = (ctypes.c_void_p * (3 * len(bns)))()
        for i, m in enumerate(bns):
            run[3 * i + 0] = m.running_mean.data_ptr()
            run[3 * i + 1] = m.running_var.data_ptr()
            run[3 * i + 2] = m.num_batches_tracked.data_ptr()
        _nt.check(lib.tg_disc_forward_groups(_nt.ptr(flat), _nt.ptr(packed), self.nb, self.ch, self.fc.in_features, _nt.ptr(x),
                                             _nt.ptr(prob), feat_ptrs, run, 1 if self.training else 0, _nt.ptr(ws),
                                             ws.numel(), n, int(groups), h, w, _nt.stream_ptr()))
        return prob, feats

    def forward_pair(self, x_real, x_fake):
        """``(self(x_real), self(x_fake))`` — the two discriminator passes of a training step (code/train.py:181,199) — as ONE
        batch: every convolution / weight-gradient kernel runs once over both halves, BatchNorm keeps separate batch
        statistics per half and updates the running statistics in call order (real, then fake), so the results are those
        of two consecutive calls at half the launches.  Returns ((prob_real, feats_real), (prob_fake, feats_fake))."""
        if x_real.shape != x_fake.shape:
            raise RuntimeError("discriminator.forward_pair: the two inputs must have the same shape")
        n = x_real.shape[0]
        prob, feats = self._forward_groups(torch.cat((x_real, x_fake), dim=0), 2)
        return (prob[:n], [f[:n] for f in feats]), (prob[n:], [f[n:] for f in feats])

    def forward(self, x):
        prob, feats = self._forward_groups(x, 1)
        return prob, feats

    def _forward_groups(self, x, groups):
        if not x.is_cuda:
            raise RuntimeError("discriminator.forward: input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 27:
            raise RuntimeError(f"discriminator.forward: expected [N,27,H,W], got {tuple(x.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if x.requires_grad:
                raise NotImplementedError("tecogan_b200 discriminator: no gradient w.r.t. the input (the reference detaches it, "
                                          "code/train.py:181,199)")
            if not self.training:
                raise RuntimeError("discriminator: backward needs train-mode BatchNorm (the reference never leaves it)")
            use_bucket, leaves = _grad_inputs(self, [p for _, p in self.named_parameters()])
            prob, *feats = _DiscriminatorFn.apply(x, self, use_bucket, groups, *leaves)
            return prob, feats
        lib = _nt.lib()
        x = x.float().contiguous()
        n, _, h, w = x.shape
        need = lib.tg_disc_workspace_bytes(n, h, w, self.nb, self.ch)
        if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        return self._run(x, self._ws, groups)


class _DiscriminatorFn(torch.autograd.Function):
    """discriminator.forward with parameter gradients (code/train.py:340).  Each call owns its workspace (the reference
    keeps two forward graphs alive, real and fake, before one backward).  Gradients enter through `prob`; the layer
    features are returned detached exactly as the reference consumes them (code/train.py:214)."""

    @staticmethod
    def forward(ctx, x, module, use_bucket, groups, *params):
        lib = _nt.lib()
        x = x.detach().float().contiguous()
        n, _, h, w = x.shape
        ws = torch.empty(lib.tg_disc_workspace_bytes(n, h, w, module.nb, module.ch), dtype=torch.uint8, device=x.device)
        prob, feats = module._run(x, ws, groups)
        ctx.module, ctx.ws, ctx.shape, ctx.use_bucket, ctx.groups = module, ws, (n, h, w), use_bucket, groups
        ctx.flat, _ = module._weights()
        ctx.packed_dgrad = module._dgrad_weights()
        ctx.save_for_backward(prob)
        ctx.mark_non_differentiable(*feats)
        return (prob, *feats)

    @staticmethod
    def backward(ctx, dprob, *dfeats):
        lib = _nt.lib()
        (prob,) = ctx.saved_tensors
        n, h, w = ctx.shape
        module = ctx.module
        params = [p for _, p in module.named_parameters()]
        if ctx.use_bucket and not _par.bucket_bound(module, params):
            raise RuntimeError("discriminator backward: the flat gradient bucket was unbound between forward and backward "
                               "(p.grad re-assigned?); re-run the forward")
        bucket = module._grad_bucket if ctx.use_bucket else None
        flat_grad = bucket if bucket is not None else torch.zeros(ctx.flat.numel(), dtype=torch.float32, device=prob.device)
        g = dprob.float().contiguous()
        _nt.check(lib.tg_disc_backward_groups(_nt.ptr(ctx.flat), _nt.ptr(ctx.packed_dgrad), module.nb, module.ch, module.fc.in_features,
                                              _nt.ptr(g), _nt.ptr(prob), _nt.ptr(flat_grad), _nt.ptr(ctx.ws), ctx.ws.numel(), n,
                                              int(ctx.groups), h, w, _nt.stream_ptr()))
        ctx.ws = None
        if bucket is not None:          # gradients were accumulated in place into the bound flat bucket (p.grad are views)
            return (None, None, None, None, None)
        grads, o = [], 0
        for p in params:
            grads.append(flat_grad[o:o + p.numel()].view_as(p).to(p.dtype))
            o += p.numel()
        return (None, None, None, None, *grads)


def f_net():
    """code/models.py:22-50 is dead code in the reference (never instantiated, main.py:231
    commented out).  Kept importable for `from models import generator, f_net, discriminator`."""
    raise NotImplementedError("f_net is unused by the reference hot path and is not provided")
