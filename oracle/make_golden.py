"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python oracle/make_golden.py
Reads /root/reference (absent on the GPU box — that is why the outputs are committed).

What runs is the reference's own code:
  * models.generator / models.discriminator / ops.upscale_four imported from
    /root/reference/code after stubbing the two absent, arithmetic-irrelevant modules
    (matplotlib, imageio — code/ops.py:9-11,20);
  * the inference frame loop is script-level code, so lines 173-219 of
    /root/reference/main.py are read from disk and exec'd with three textual shims:
    ``.cuda()`` and ``.cpu()`` removed, ``.half()`` -> ``.half().float()`` (CPU grid_sample
    rejects mixed dtypes).  Nothing is copied into this repository.
Weights/inputs come from oracle/synth.py so they can be regenerated anywhere.
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    for name in ["matplotlib", "matplotlib.animation", "matplotlib.pyplot", "imageio"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.animation"].ArtistAnimation = object
    sys.modules["matplotlib.animation"].PillowWriter = object
    sys.path.insert(1, os.path.join(REF, "code"))
    import models as ref_models  # noqa
    import ops as ref_ops  # noqa
    return ref_models, ref_ops


def run_reference_loop(ref_models, ref_ops, G, r_inputs, crop):
    """exec main.py:173-219 (the per-clip body of the inference loop)."""
    src = open(os.path.join(REF, "main.py")).read().split("\n")
    body = "\n".join(src[174:219])          # lines 175..219: body of `for batch_idx, r_inputs`
    body = textwrap.dedent(body)
    body = body.replace(".cuda()", "").replace(".cpu()", "").replace(".half()", ".half().float()")
    ns = dict(vars(ref_ops))
    ns.update(torch=torch, F=torch.nn.functional, generator_F=G, r_inputs=r_inputs,
              args=types.SimpleNamespace(crop_size=crop, learning_rate=1e-4))
    with torch.no_grad():
        exec(compile(body, "reference_main_loop", "exec"), ns)
    return ns["gen_outputs"]


def grad_fingerprint(module):
    """per-parameter (norm, projection on a fixed PCG64 direction) of .grad — a small, order-sensitive fingerprint."""
    from oracle import synth
    rows = []
    for i, (_, p) in enumerate(module.named_parameters()):
        g = p.grad.detach().double().flatten()
        d = torch.from_numpy(synth.det_uniform((g.numel(),), 9000 + i, -1.0, 1.0)).double()
        rows.append((float(g.norm()), float(g @ d)))
    return np.array(rows, np.float64)


def write_train_golden(ref_models, out_dir, name="train.npz", crop=32, sub=8, **flags):
    """One step of the UNMODIFIED reference train.FRVSR_Train (code/train.py:374-377 -> TecoGAN, :49-370) on CPU.
    Shims: torch.Tensor.cuda -> identity; F.grid_sample casts the grid to the image dtype (CPU grid_sample
    rejects the fp16 grid of train.py:98,187; the fp16 rounding itself is kept).  GradScaler / autocast disable
    themselves without CUDA, so the step is plain fp32.

    crop=64 is BASELINE cfg5's shape (256x256 HR): the reference discriminator hard-codes fc = denselayer(48, 1)
    (code/models.py:123) and colab/README.md:15-22 tells users to edit that line to denselayer(192, 1); the same edit is
    applied here to the constructed module (nothing else of the reference changes).  **flags override argparse defaults
    (e.g. pingpang=True, code/train.py:56-62,153-156,275-285)."""
    import warnings
    from oracle import synth, train_oracle
    F = torch.nn.functional
    warnings.filterwarnings("ignore")
    orig_cuda, orig_gs = torch.Tensor.cuda, F.grid_sample
    torch.Tensor.cuda = lambda self, *a, **k: self
    F.grid_sample = lambda inp, grid, *a, **k: orig_gs(inp, grid.to(inp.dtype), *a, **k)
    try:
        import train as ref_train
        import ops as ref_ops
        args = train_oracle.default_train_args(crop_size=crop, **flags)
        G = ref_models.generator(3, args=args)
        D = ref_models.discriminator(args=args)
        if crop != 32:
            D.fc = ref_ops.denselayer(48 * (crop // 32) ** 2, 1)          # colab/README.md:15-22
        G.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fill_state_dict(G.state_dict(), seed=1, gain=1.0).items()})
        D.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fill_state_dict(D.state_dict(), seed=2, gain=1.0).items()})
        og = torch.optim.Adam(G.parameters(), args.learning_rate, betas=(args.beta, 0.999), eps=args.adameps)   # main.py:239-243
        od = torch.optim.Adam(D.parameters(), args.learning_rate, betas=(args.beta, 0.999), eps=args.adameps)
        b = 2
        r_in = torch.from_numpy(synth.det_uniform((b, 10, 3, crop, crop), 51, 0.0, 1.0))
        r_tg = torch.from_numpy(synth.det_uniform((b, 10, 3, 4 * crop, 4 * crop), 52, 0.0, 1.0))
        w0 = G.conv[0].weight.detach().clone()
        out = ref_train.FRVSR_Train(r_in, r_tg, args, D, G, 0, 0.0, 0.0, og, od)
        n = len(out.update_list)
        np.savez_compressed(
            os.path.join(out_dir, name),
            names=np.array(out.update_list_name[:n]),
            update_list=np.array([float(v) for v in out.update_list], np.float64),
            update_list_avg=np.array([float(v) for v in out.update_list_avg[:n]], np.float64),
            tb=np.float64(float(out.tb)), dt_ratio=np.float64(float(out.update_list_avg[n + 1])),
            d_loss=np.float64(float(out.d_loss)), gen_loss=np.float64(float(out.gen_loss)),
            gen_output_sub=out.gen_output.detach()[:, :, :, ::sub, ::sub].numpy().astype(np.float32),
            target_sub=out.target.detach()[:, :, ::sub, ::sub].numpy().astype(np.float32),
            g_grad=grad_fingerprint(G), d_grad=grad_fingerprint(D),
            g_conv0_step=(G.conv[0].weight.detach() - w0)[:4, :4].numpy(),
            d_running_mean_block1=D.block1[1].running_mean.numpy(),
            batch=b, crop=crop)
    finally:
        torch.Tensor.cuda, F.grid_sample = orig_cuda, orig_gs


def main():
    from oracle import synth
    ref_models, ref_ops = import_reference()
    torch.set_num_threads(4)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    args = types.SimpleNamespace(num_resblock=16, discrim_resblocks=4, discrim_channels=128)

    # ---- glue ops, exactly the calls the reference makes -------------------------------
    F = torch.nn.functional
    img = torch.from_numpy(synth.det_uniform((2, 3, 16, 24), 11, -1.0, 1.0))
    grid = torch.from_numpy(synth.det_uniform((2, 16, 24, 2), 12, -1.2, 1.2))
    lr = torch.from_numpy(synth.det_uniform((2, 3, 6, 10), 13, 0.0, 1.0))
    np.savez_compressed(
        os.path.join(out_dir, "glue.npz"),
        warp=F.grid_sample(img, grid.half().float()).numpy(),                   # main.py:203
        upscale=ref_ops.upscale_four(lr * 4.0).numpy(),                         # main.py:186
        s2d=img.view(2, 3, 4, 4, 6, 4).permute(0, 1, 3, 5, 2, 4).reshape(2, 48, 4, 6).numpy(),  # :207-212
        deprocess=ref_ops.deprocess(img).numpy(), preprocess=ref_ops.preprocess(img).numpy(),
        torch_version=str(torch.__version__))

    # ---- generator: single forward + the full recurrent loop ---------------------------
    for tag, gain in (("g1", 1.0), ("g17", 1.7)):
        G = ref_models.generator(3, args=args).eval()
        named = synth.fill_state_dict(G.state_dict(), seed=1, gain=gain)
        G.load_state_dict({k: torch.from_numpy(v) for k, v in named.items()})
        crop, T = 16, 4
        r_inputs = torch.from_numpy(synth.clip_inputs(1, T, crop, crop, seed=1234, hi=0.25))
        outs = run_reference_loop(ref_models, ref_ops, G, r_inputs, crop)        # [T,3,64,64]
        x51 = torch.from_numpy(synth.det_uniform((1, 51, 12, 20), 21, 0.0, 1.0))
        with torch.no_grad():
            y = G(x51)
        np.savez_compressed(os.path.join(out_dir, f"gen_{tag}.npz"),
                            loop_out=outs.numpy().astype(np.float32),
                            fwd_out=y.numpy().astype(np.float32),
                            gain=np.float32(gain), crop=crop, T=T)

    # ---- discriminator forward (train-mode BN, as the reference always runs it) --------
    D = ref_models.discriminator(args=args)
    named = synth.fill_state_dict(D.state_dict(), seed=2, gain=1.0)
    D.load_state_dict({k: torch.from_numpy(v) for k, v in named.items()})
    D.train()
    x = torch.from_numpy(synth.det_uniform((3, 27, 128, 128), 31, -1.0, 1.0))
    with torch.no_grad():
        prob, feats = D(x)
    np.savez_compressed(os.path.join(out_dir, "disc.npz"), prob=prob.numpy(),
                        f4=feats[3].numpy(),
                        f_abs_mean=np.array([f.abs().mean().item() for f in feats], np.float64),
                        f_mean=np.array([f.mean().item() for f in feats], np.float64),
                        running_mean_block1=D.block1[1].running_mean.numpy())
    write_train_golden(ref_models, out_dir)
    write_train_golden(ref_models, out_dir, name="train_cfg5.npz", crop=64, sub=16)        # BASELINE cfg5 shape
    write_train_golden(ref_models, out_dir, name="train_pingpang.npz", pingpang=True)       # code/train.py:56-62,153-156,275-285
    print("golden fixtures written to", out_dir)
    for f in sorted(os.listdir(out_dir)):
        print(" ", f, os.path.getsize(os.path.join(out_dir, f)), "bytes")


if __name__ == "__main__":
    main()
