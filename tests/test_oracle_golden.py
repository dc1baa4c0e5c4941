"""CPU: pin the oracle (oracle/) to outputs of the unmodified reference (tests/golden/*.npz,
written by oracle/make_golden.py from /root/reference)."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import glue_np, synth, tecogan_oracle as O

warnings.filterwarnings("ignore", category=UserWarning)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_glue_numpy_vs_reference(golden_dir):
    g = _load(golden_dir, "glue.npz")
    img = synth.det_uniform((2, 3, 16, 24), 11, -1.0, 1.0)
    grid = synth.det_uniform((2, 16, 24, 2), 12, -1.2, 1.2)
    lr = synth.det_uniform((2, 3, 6, 10), 13, 0.0, 1.0)
    assert np.array_equal(glue_np.space_to_depth(img, 4), g["s2d"])                 # bit-exact
    assert np.array_equal(glue_np.depth_to_space(g["s2d"], 4), img)                 # round trip
    np.testing.assert_allclose(glue_np.warp(img, grid), g["warp"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(glue_np.upscale_four(lr * np.float32(4.0)), g["upscale"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(glue_np.deprocess(img), g["deprocess"], rtol=0, atol=0)
    np.testing.assert_allclose(glue_np.preprocess(img), g["preprocess"], rtol=0, atol=0)


def test_glue_torch_oracle_vs_reference(golden_dir):
    g = _load(golden_dir, "glue.npz")
    img = torch.from_numpy(synth.det_uniform((2, 3, 16, 24), 11, -1.0, 1.0))
    grid = torch.from_numpy(synth.det_uniform((2, 16, 24, 2), 12, -1.2, 1.2))
    lr = torch.from_numpy(synth.det_uniform((2, 3, 6, 10), 13, 0.0, 1.0))
    assert np.array_equal(O.space_to_depth(img, 4).numpy(), g["s2d"])
    assert np.array_equal(O.depth_to_space(torch.from_numpy(g["s2d"]), 4).numpy(), img.numpy())
    np.testing.assert_allclose(O.warp(img, grid).numpy(), g["warp"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(O.upscale_four(lr * 4.0).numpy(), g["upscale"], rtol=0, atol=1e-7)


def test_frame_input_numpy_vs_torch():
    lr_t = synth.det_uniform((2, 3, 5, 7), 41, 0, 0.25)
    lr_p = synth.det_uniform((2, 3, 5, 7), 42, 0, 0.25)
    hr = synth.det_uniform((2, 3, 20, 28), 43, 0, 1)
    a = glue_np.frame_input(lr_t, lr_p, hr)
    b = O.frame_input(torch.from_numpy(lr_t), torch.from_numpy(lr_p), torch.from_numpy(hr)).numpy()
    # a 1-ulp fp32 difference in the upscale can flip an fp16 grid ulp: allow
    # a handful of outliers, everything else must agree to 1e-5
    bad = np.abs(a - b) > 1e-5
    assert bad.mean() < 2e-3, bad.mean()


@pytest.mark.parametrize("tag,gain", [("g1", 1.0), ("g17", 1.7)])
def test_generator_and_loop_vs_reference(golden_dir, tag, gain):
    g = _load(golden_dir, f"gen_{tag}.npz")
    torch.set_num_threads(4)
    G = O.OracleGenerator(3, 16).eval()
    O.load_numpy_state(G, synth.fill_state_dict(G.state_dict(), seed=1, gain=gain))
    x51 = torch.from_numpy(synth.det_uniform((1, 51, 12, 20), 21, 0.0, 1.0))
    with torch.no_grad():
        y = G(x51).numpy()
    np.testing.assert_allclose(y, g["fwd_out"], rtol=0, atol=2e-6)
    crop, T = int(g["crop"]), int(g["T"])
    r = torch.from_numpy(synth.clip_inputs(1, T, crop, crop, seed=1234, hi=0.25))
    out = O.infer_clip(G, r)[0].numpy()
    assert out.shape == g["loop_out"].shape
    np.testing.assert_allclose(out, g["loop_out"], rtol=0, atol=5e-6)


def test_discriminator_vs_reference(golden_dir):
    g = _load(golden_dir, "disc.npz")
    torch.set_num_threads(4)
    D = O.OracleDiscriminator(4, 128, 48)
    O.load_numpy_state(D, synth.fill_state_dict(D.state_dict(), seed=2, gain=1.0))
    D.train()
    x = torch.from_numpy(synth.det_uniform((3, 27, 128, 128), 31, -1.0, 1.0))
    with torch.no_grad():
        prob, feats = D(x)
    np.testing.assert_allclose(prob.numpy(), g["prob"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(feats[3].numpy(), g["f4"], rtol=0, atol=1e-4)
    np.testing.assert_allclose([f.abs().mean().item() for f in feats], g["f_abs_mean"], rtol=1e-5)
    np.testing.assert_allclose(D.block1[1].running_mean.numpy(), g["running_mean_block1"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("fixture,crop,flags,sub", [("train.npz", 32, {}, 8), ("train_cfg5.npz", 64, {}, 16),
                                                    ("train_pingpang.npz", 32, {"pingpang": True}, 8)],
                         ids=["cfg4_shape", "cfg5_shape", "pingpang"])
def test_train_step_oracle_vs_reference(golden_dir, fixture, crop, flags, sub):
    """oracle/train_oracle.py against one step of the unmodified reference train.FRVSR_Train (tests/golden/train*.npz):
    every logged scalar, the EMA list, the generator outputs, the discriminator's real input, both nets' gradients
    (per-tensor norm + projection fingerprint) and the Adam update — at BASELINE cfg4's crop (32), cfg5's crop (64, the
    reference discriminator with fc = denselayer(192, 1) as colab/README.md:15-22 prescribes) and with pingpang=True."""
    from oracle import train_oracle as TO
    g = _load(golden_dir, fixture)
    torch.set_num_threads(8)
    args = TO.default_train_args(crop_size=crop, **flags)
    G = O.OracleGenerator(3, 16)
    D = O.OracleDiscriminator(4, 128, 48 * (crop // 32) ** 2)
    O.load_numpy_state(G, synth.fill_state_dict(G.state_dict(), seed=1, gain=1.0))
    O.load_numpy_state(D, synth.fill_state_dict(D.state_dict(), seed=2, gain=1.0))
    og = torch.optim.Adam(G.parameters(), args.learning_rate, betas=(args.beta, 0.999), eps=args.adameps)
    od = torch.optim.Adam(D.parameters(), args.learning_rate, betas=(args.beta, 0.999), eps=args.adameps)
    b = int(g["batch"])
    r_in = torch.from_numpy(synth.det_uniform((b, 10, 3, crop, crop), 51, 0.0, 1.0))
    r_tg = torch.from_numpy(synth.det_uniform((b, 10, 3, 4 * crop, 4 * crop), 52, 0.0, 1.0))
    w0 = G.conv[0].weight.detach().clone()
    out = TO.train_step(G, D, og, od, r_in, r_tg, args, 0)
    assert list(out["log"].keys()) == [str(n) for n in g["names"]]
    np.testing.assert_allclose(list(out["log"].values()), g["update_list"], rtol=2e-5)
    np.testing.assert_allclose(out["log_avg"], g["update_list_avg"], rtol=2e-5)
    np.testing.assert_allclose([out["tb"], out["dt_ratio"], out["d_loss"], out["gen_loss"]],
                               [g["tb"], g["dt_ratio"], g["d_loss"], g["gen_loss"]], rtol=2e-5)
    np.testing.assert_allclose(out["gen_output"][:, :, :, ::sub, ::sub].numpy(), g["gen_output_sub"], atol=2e-6)
    np.testing.assert_allclose(out["target"][:, :, ::sub, ::sub].numpy(), g["target_sub"], atol=2e-6)
    for mod, key in ((G, "g_grad"), (D, "d_grad")):
        for i, (name, p) in enumerate(mod.named_parameters()):
            gr = p.grad.detach().double().flatten()
            d = torch.from_numpy(synth.det_uniform((gr.numel(),), 9000 + i, -1.0, 1.0)).double()
            want_norm, want_proj = g[key][i]
            assert abs(float(gr.norm()) - want_norm) <= 3e-3 * want_norm + 1e-9, (name, float(gr.norm()), want_norm)
            # fp32 summation order (thread count, oneDNN blocking) moves the discriminator's BatchNorm-coupled gradients by
            # a few 1e-3 of their norm between runs of the reference itself
            assert abs(float(gr @ d) - want_proj) <= 1e-2 * want_norm + 1e-9, (name, float(gr @ d), want_proj)
    np.testing.assert_allclose((G.conv[0].weight.detach() - w0)[:4, :4].numpy(), g["g_conv0_step"], atol=2e-6)
    np.testing.assert_allclose(D.block1[1].running_mean.numpy(), g["d_running_mean_block1"], atol=1e-6)
