"""Fused optimizer step for the training mirror.

The reference's training driver builds two stock optimizers, ``torch.optim.Adam(net.parameters(), lr, betas=(beta, 0.999),
eps=adameps)`` (main.py:239-243), and ``TecoGAN`` steps them under one module-global ``GradScaler``
(code/train.py:9,335-342).  ``FlatAdam.adopt(module, optimizer)`` keeps that optimizer object as the container — its
``param_groups`` stay the source of lr / betas / eps (StepLR keeps working, main.py:247-248,296-297), its ``state`` holds
``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter so ``optimizer.state_dict()`` checkpoints as before
(main.py:308-317) — but

* every parameter of the module becomes a view of ONE flat f32 buffer (same layout as the flat gradient bucket of
  tecogan_b200.parallel and as the flat parameter vector tg_gen_pack / tg_disc_pack read), and so do the Adam moments;
* the step is three launches of repo kernels (include/tecogan_b200.h: tg_grad_check_finite, tg_adam_step,
  tg_scaler_update) + the batched bf16 re-pack of the updated weights, instead of ~100 multi-tensor launches and a
  gather copy; lr, step count, loss scale and the inf/NaN flag live in device scalars, so the whole training step can be
  captured in a CUDA graph (tecogan_b200.train) and still follow StepLR and the GradScaler.

Optimizers that are not plain Adam (amsgrad, weight decay, several param groups, foreign parameters) are not adopted:
``FlatAdam.adopt`` returns None and the training step falls back to ``scaler.step(optimizer)``.
"""
import torch

from . import _native as _nt


def bind_flat_params(module):
    """Make every parameter of ``module`` a view of one flat f32 device buffer (named_parameters order).  Idempotent;
    re-binds after ``module.to(...)`` / ``.cuda()`` replaced the storages.  Returns the flat buffer."""
    params = [p for _, p in module.named_parameters()]
    total = sum(p.numel() for p in params)
    flat = getattr(module, "_flat_params", None)
    ok = flat is not None and flat.numel() == total and flat.device == params[0].device
    if ok:
        o = flat.data_ptr()
        for p in params:
            if p.data_ptr() != o or p.dtype != torch.float32:
                ok = False
                break
            o += 4 * p.numel()
    if ok:
        return flat
    dev = params[0].device
    if dev.type != "cuda":
        raise RuntimeError("bind_flat_params: parameters must live on a CUDA device (call .cuda() first)")
    flat = torch.empty(total, dtype=torch.float32, device=dev)
    o = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            flat[o:o + n].copy_(p.detach().reshape(-1).float())
            p.data = flat[o:o + n].view(p.shape)
            o += n
    module._flat_params = flat
    if hasattr(module, "invalidate_packed"):
        module.invalidate_packed()
    return flat


class FlatAdam:
    """The Adam state of one (module, torch.optim.Adam) pair on flat buffers + its device scalars."""

    @staticmethod
    def adoptable(module, optimizer):
        if type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
            return False
        g = optimizer.param_groups[0]
        if g.get("amsgrad") or g.get("maximize") or g.get("weight_decay", 0) != 0 or g.get("differentiable"):
            return False
        if isinstance(g["lr"], torch.Tensor) and g["lr"].numel() != 1:
            return False
        mine = [p for _, p in module.named_parameters()]
        theirs = g["params"]
        return len(mine) == len(theirs) and all(a is b for a, b in zip(mine, theirs)) and all(p.is_cuda for p in mine)

    @classmethod
    def adopt(cls, module, optimizer):
        """FlatAdam for this pair (cached on the optimizer object), or None when the optimizer is not plain Adam over
        exactly this module's parameters."""
        fa = getattr(optimizer, "_tg_flat_adam", None)
        if fa is not None and fa.module is module:
            fa.sync()
            return fa
        if not cls.adoptable(module, optimizer):
            return None
        fa = cls(module, optimizer)
        optimizer._tg_flat_adam = fa
        return fa

    def __init__(self, module, optimizer):
        self.module, self.opt = module, optimizer
        self.params = [p for _, p in module.named_parameters()]
        self.flat = bind_flat_params(module)
        dev = self.flat.device
        n = self.flat.numel()
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step = torch.zeros((), dtype=torch.float32, device=dev)        # optimizer.state[p]['step'] for every p
        self.lr = torch.zeros((), dtype=torch.float32, device=dev)
        self.found_inf = torch.zeros((), dtype=torch.float32, device=dev)
        self._lr_host = None
        self._register(copy_from_state=True)

    # -------------------------------------------------------------------------------------- state <-> optimizer.state
    def _register(self, copy_from_state):
        """Point optimizer.state[p] at views of the flat moments (taking over whatever state the optimizer already holds,
        e.g. after optimizer.load_state_dict)."""
        st = self.opt.state
        o = 0
        step_val = None
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                ea, es = self.exp_avg[o:o + n].view(p.shape), self.exp_avg_sq[o:o + n].view(p.shape)
                old = st.get(p)
                if copy_from_state and old:
                    if "exp_avg" in old:
                        ea.copy_(old["exp_avg"])
                        es.copy_(old["exp_avg_sq"])
                    if "step" in old:
                        step_val = float(old["step"])
                st[p] = {"step": self.step, "exp_avg": ea, "exp_avg_sq": es}
                o += n
            if step_val is not None:
                self.step.fill_(step_val)
        self._state_key = self._key()

    def _key(self):
        st = self.opt.state
        p0, p1 = self.params[0], self.params[-1]
        return (st[p0]["exp_avg"].data_ptr() if p0 in st and "exp_avg" in st[p0] else -1,
                st[p1]["exp_avg"].data_ptr() if p1 in st and "exp_avg" in st[p1] else -1, self.flat.data_ptr())

    def sync(self):
        """Re-adopt after the caller replaced state or storages (optimizer.load_state_dict, module.to(...))."""
        flat = bind_flat_params(self.module)
        if flat is not self.flat:
            self.flat = flat
        if self._key() != self._state_key:
            self._register(copy_from_state=True)

    def refresh_lr(self):
        """lr of the param group -> device scalar (one tiny fill when StepLR changed it; never inside a graph)."""
        lr = self.opt.param_groups[0]["lr"]
        lr = float(lr)
        if lr != self._lr_host:
            self.lr.fill_(lr)
            self._lr_host = lr

    # ------------------------------------------------------------------------------------------------------ the step
    def step_kernels(self, grads, inv_scale=None, scale=None, growth_tracker=None, growth=2.0, backoff=0.5, interval=2000):
        """check inf/NaN -> Adam -> scaler / step-counter update -> bf16 re-pack of the updated weights.  `grads` is the
        module's flat gradient bucket (scaled by the loss scale when inv_scale is given).  No host synchronisation."""
        lib = _nt.lib()
        g = self.opt.param_groups[0]
        b1, b2 = g["betas"]
        st = _nt.stream_ptr(self.flat.device)
        n = self.flat.numel()
        if grads.numel() != n or grads.dtype != torch.float32:
            raise RuntimeError("FlatAdam.step_kernels: gradient bucket does not match the flat parameter buffer")
        _nt.check(lib.tg_grad_check_finite(_nt.ptr(grads), n, _nt.ptr(self.found_inf), st))
        _nt.check(lib.tg_adam_step(_nt.ptr(self.flat), _nt.ptr(grads), _nt.ptr(self.exp_avg), _nt.ptr(self.exp_avg_sq), n,
                                   _nt.ptr(self.lr), float(b1), float(b2), float(g["eps"]), _nt.ptr(self.step),
                                   _nt.ptr(inv_scale), _nt.ptr(self.found_inf), st))
        _nt.check(lib.tg_scaler_update(_nt.ptr(scale), _nt.ptr(growth_tracker), _nt.ptr(self.found_inf), float(growth),
                                       float(backoff), int(interval), _nt.ptr(self.step), st))
        self.module.repack_from_flat()
