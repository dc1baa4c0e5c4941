"""GPU parity of the fused flat-bucket Adam (tecogan_b200.optim) against torch.optim.Adam — the optimizer the
reference builds (main.py:239-243) and steps under a GradScaler (code/train.py:335-342) — and of the CUDA-graphed training
step against the eager one."""
import copy
import types

import pytest
import torch

from oracle import synth, train_oracle as TO

pytestmark = pytest.mark.gpu


def _gen(nres=1):
    from tecogan_b200 import models
    torch.manual_seed(3)
    return models.generator(3, types.SimpleNamespace(num_resblock=nres)).cuda()


def _adam(m, lr=1e-4):
    return torch.optim.Adam(m.parameters(), lr, betas=(0.9, 0.999), eps=1e-8)          # main.py:239-243


def test_flat_adam_matches_torch_adam_and_keeps_the_optimizer_contract():
    from tecogan_b200 import optim, parallel
    G, Gr = _gen(), _gen()
    Gr.load_state_dict(G.state_dict())
    og, ogr = _adam(G), _adam(Gr)
    fa = optim.FlatAdam.adopt(G, og)
    assert fa is not None and optim.FlatAdam.adopt(G, og) is fa
    # parameters are views of one flat buffer, module API unchanged
    assert all(p.data_ptr() >= fa.flat.data_ptr() for p in G.parameters())
    sd0 = {k: v.clone() for k, v in G.state_dict().items()}
    bucket = parallel.bind_flat_grads(G)
    sched = torch.optim.lr_scheduler.StepLR(og, 2, 0.5)                               # main.py:247-248
    sched_r = torch.optim.lr_scheduler.StepLR(ogr, 2, 0.5)
    for it in range(4):
        o = 0
        for i, (p, pr) in enumerate(zip(G.parameters(), Gr.parameters())):
            g = torch.from_numpy(synth.det_uniform(tuple(p.shape), 100 * it + i, -1.0, 1.0)).cuda() * (10.0 ** (i % 5 - 3))
            bucket[o:o + p.numel()].copy_(g.reshape(-1))
            pr.grad = g.clone()
            o += p.numel()
        fa.refresh_lr()
        fa.step_kernels(bucket)
        ogr.step()
        sched.step()
        sched_r.step()
    for (name, p), pr in zip(G.named_parameters(), Gr.parameters()):
        assert torch.allclose(p, pr, rtol=1e-5, atol=1e-7), name
        assert (p - sd0[name]).abs().max().item() > 0
        st, str_ = og.state[p], ogr.state[pr]
        # torch forms exp_avg with lerp_ (m + (g - m) * (1 - b1)), the kernel with b1 * m + (1 - b1) * g: equal to rounding
        for k in ("exp_avg", "exp_avg_sq"):
            assert (st[k] - str_[k]).abs().max().item() <= 1e-4 * str_[k].abs().max().item(), (name, k)
        assert float(st["step"]) == float(str_["step"]) == 4.0
    # checkpoint round trip through the stock optimizer API (main.py:308-317, 252-258): state_dict -> a fresh torch Adam
    fresh = _adam(Gr)
    fresh.load_state_dict(copy.deepcopy(og.state_dict()))
    assert float(fresh.state[next(iter(Gr.parameters()))]["step"]) == 4.0
    # ... and loading a state dict INTO the adopted optimizer is picked up at the next adopt()
    og.load_state_dict(copy.deepcopy(ogr.state_dict()))
    fa2 = optim.FlatAdam.adopt(G, og)
    assert fa2 is fa and float(fa.step) == 4.0
    p0 = next(iter(G.parameters()))
    assert og.state[p0]["exp_avg"].data_ptr() >= fa.exp_avg.data_ptr()
    # the packed bf16 weights follow the update without an explicit invalidate
    x = torch.from_numpy(synth.det_uniform((1, 51, 16, 16), 9, 0.0, 1.0)).cuda()
    with torch.no_grad():
        G.eval(); Gr.eval()
        assert torch.allclose(G(x), Gr(x), atol=2e-3)


def test_flat_adam_skips_on_inf_and_drives_the_grad_scaler():
    from tecogan_b200 import optim, parallel
    G = _gen()
    og = _adam(G)
    fa = optim.FlatAdam.adopt(G, og)
    bucket = parallel.bind_flat_grads(G)
    sc = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=2)
    one = torch.ones((), device="cuda")
    s0 = sc.scale(one).clone()
    before = fa.flat.clone()
    bucket.fill_(1024.0)
    bucket[5] = float("inf")
    fa.refresh_lr()
    kw = dict(scale=sc._scale, growth_tracker=sc._growth_tracker, growth=sc.get_growth_factor(), backoff=sc.get_backoff_factor(),
              interval=sc.get_growth_interval())
    fa.step_kernels(bucket, torch.reciprocal(s0), **kw)
    assert torch.equal(fa.flat, before) and float(fa.step) == 0.0 and float(fa.found_inf) == 1.0
    assert sc.get_scale() == 512.0                                    # backed off
    for k in range(2):                                                # two clean steps -> the scale grows back
        bucket.fill_(512.0 * 3.0)
        fa.step_kernels(bucket, torch.reciprocal(sc.scale(one)).clone(), **kw)
        assert float(fa.found_inf) == 0.0 and float(fa.step) == k + 1.0
        assert torch.allclose(bucket, torch.full_like(bucket, 3.0))   # unscaled in place, like GradScaler.unscale_
    assert sc.get_scale() == 1024.0
    assert (fa.flat - before).abs().max().item() > 0


def test_not_adoptable_optimizers_fall_back():
    from tecogan_b200 import optim
    G = _gen()
    assert optim.FlatAdam.adopt(G, torch.optim.Adam(G.parameters(), 1e-4, weight_decay=1e-2)) is None
    assert optim.FlatAdam.adopt(G, torch.optim.SGD(G.parameters(), 1e-4)) is None
    assert optim.FlatAdam.adopt(G, torch.optim.Adam(list(G.parameters())[:-1], 1e-4)) is None


def test_graphed_train_step_matches_eager():
    """Five FRVSR_Train steps on a fixed batch: with CUDA-graph capture (steps 3-5 are replays of one captured graph) and
    fully eager, from identical initial states.  The wgrad kernels accumulate with f32 atomics, so two runs agree to
    rounding, not bit for bit, and Adam's first steps (+-lr * sign(g)) amplify that where a gradient is ~0: every logged
    scalar within 1 % after five steps (measured 0.2 %), final parameters within a few Adam steps' noise."""
    from tecogan_b200 import models, train as T
    args = TO.default_train_args(num_resblock=2, discrim_resblocks=1, discrim_channels=64)
    r_in = torch.from_numpy(synth.det_uniform((2, 10, 3, 32, 32), 71, 0.0, 1.0)).cuda()
    r_tg = torch.from_numpy(synth.det_uniform((2, 10, 3, 128, 128), 72, 0.0, 1.0)).cuda()
    runs = {}
    saved = (T.USE_CUDA_GRAPH, T.scaler)
    try:
        for mode in (True, False):
            T.USE_CUDA_GRAPH = mode
            T.scaler = None                                            # a fresh module-global GradScaler per run
            T._graphs.clear()
            torch.manual_seed(11)
            G, D = models.generator(3, args).cuda(), models.discriminator(args).cuda()
            og, od = _adam(G), _adam(D)
            logs = []
            for step in range(5):
                # `out` of the previous step stays alive across the call, as in main.py's loop (:274-282): its autograd
                # graph keeps the parameters' AccumulateGrad nodes of an EAGER step alive while the next step is captured
                out = T.FRVSR_Train(r_in, r_tg, args, D, G, step, 0.0, 0.0, og, od)
                assert out.global_step == step + 1
                content = float(torch.mean(torch.sum(torch.square(out.gen_output.detach() - r_tg), dim=[4])))
                logs.append([float(v) for v in out.update_list] + [float(out.d_loss), content])
            if mode:
                gs = list(T._graphs.values())[-1]
                assert gs.graph is not None and gs.calls == 5          # steps 3..5 really were graph replays
            runs[mode] = (logs, torch.cat([p.detach().flatten() for p in G.parameters()]).clone(),
                          torch.cat([p.detach().flatten() for p in D.parameters()]).clone(), float(og.state[next(iter(G.parameters()))]["step"]),
                          int(D.block1[1].num_batches_tracked))
        key = [k for k in T._graphs]
    finally:
        T.USE_CUDA_GRAPH, T.scaler = saved
        T._graphs.clear()
    (lg, pg, pd, sg, ng), (le, pe, pde, se, ne) = runs[True], runs[False]
    assert sg == se == 5.0 and ng == ne == 10
    for a, b in zip(lg, le):
        for x, y in zip(a, b):
            assert abs(x - y) <= 1e-2 * max(1.0, abs(y)), (a, b)
    assert lg[4][-1] < lg[0][-1] and le[4][-1] < le[0][-1]              # the pure content loss decreases on the fixed batch
    assert (pg - pe).abs().max().item() <= 1.2e-3 and (pg - pe).abs().mean().item() <= 5e-5      # 5 steps of lr = 1e-4
    assert (pd - pde).abs().max().item() <= 1.2e-3
