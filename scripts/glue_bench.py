"""HBM roofline of the memory-bound glue kernels at BASELINE cfg3 sizes (960x540 LR -> 3840x2160 HR), through the C ABI.

Algorithmic bytes per unit are SURVEY.md section 8(d)'s: warp 28 B per HR pixel (3 f32 in, 3 f32 out, fp16-rounded
2-channel grid read as f32 = 32 B actually moved), space_to_depth / depth_to_space 24 B, upscale_four 12.75 B, the fused
frame-input producer 18.4 B.  Buffers (>= 0.4 GB per call) are far larger than the 126 MB L2; peak = MEASURED_PEAKS.json
hbm_gbs.  Prints one JSON line; bench.py embeds the same measurement as its "glue" object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-tecogan_b200"))


def measure(n=2, h=540, w=960, iters=10, dev=None):
    from tecogan_b200 import _native as nt
    if os.environ.get("TG_GLUE_LIB"):                         # measurement only: another build of the library (same-box A/B)
        nt.LIB_PATH = os.path.abspath(os.environ["TG_GLUE_LIB"])
    lib = nt.lib()
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    ho, wo = 4 * h, 4 * w
    hr_px = n * ho * wo
    g = torch.Generator(device=dev).manual_seed(7)
    hr = torch.rand((n, 3, ho, wo), device=dev, generator=g)
    hr2 = torch.empty_like(hr)
    lr = torch.rand((n, 3, h, w), device=dev, generator=g) * 0.25
    lr_prev = torch.rand((n, 3, h, w), device=dev, generator=g) * 0.25       # flow in [0,1): every warp tap in bounds
    grid = (torch.rand((n, ho, wo, 2), device=dev, generator=g) * 2 - 1) * 0.98         # every pixel samples a random place
    # a jittered field: the identity sampling grid plus an INDEPENDENT displacement of up to +-2 pixels per pixel (neighbouring
    # pixels sample up to five different rows: harsher than any real flow, kept for continuity with round 1)
    ys = ((torch.arange(ho, device=dev, dtype=torch.float32) + 0.5) * 2 / ho - 1).view(1, ho, 1).expand(n, ho, wo)
    xs = ((torch.arange(wo, device=dev, dtype=torch.float32) + 0.5) * 2 / wo - 1).view(1, 1, wo).expand(n, ho, wo)
    smooth = torch.stack((xs, ys), dim=-1) + (torch.rand((n, ho, wo, 2), device=dev, generator=g) - 0.5) * (8.0 / wo)
    smooth = smooth.contiguous()
    # what upscale_four(flow) hands the reference's warp (code/train.py:94-101): a spatially smooth displacement (here up
    # to +-2 pixels, varying over ~100-pixel periods), so neighbouring pixels sample neighbouring taps
    yy = torch.arange(ho, device=dev, dtype=torch.float32).view(1, ho, 1)
    xx = torch.arange(wo, device=dev, dtype=torch.float32).view(1, 1, wo)
    dxs = 2.0 * torch.sin(yy * (6.2831853 / 97.0)) * torch.cos(xx * (6.2831853 / 131.0))
    dys = 2.0 * torch.cos(yy * (6.2831853 / 113.0)) * torch.sin(xx * (6.2831853 / 89.0))
    flowlike = torch.stack((xs + dxs * (2.0 / wo), ys + dys * (2.0 / ho)), dim=-1).contiguous()
    depth = torch.empty((n, 48, h, w), device=dev)
    x_in = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=dev)
    st = nt.stream_ptr()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / iters

    big = torch.empty(hr.numel() * 2, device=dev)             # reference lines on the same box: a copy and a fill
    big2 = torch.empty_like(big)
    ops = {
        "reference_copy_f32 (read + write)": (8.0 * big.numel() / hr_px, lambda: big2.copy_(big)),
        "reference_fill_f32 (write only)": (4.0 * big.numel() / hr_px, lambda: big.zero_()),
        "space_to_depth": (24.0, lambda: nt.check(lib.tg_space_to_depth(nt.ptr(hr), nt.ptr(depth), n, 3, h, w, 4, st))),
        "depth_to_space": (24.0, lambda: nt.check(lib.tg_depth_to_space(nt.ptr(depth), nt.ptr(hr2), n, 3, h, w, 4, st))),
        "warp_bilinear_smooth_field": (28.0, lambda: nt.check(lib.tg_warp_bilinear(nt.ptr(hr), nt.ptr(flowlike), nt.ptr(hr2), n, 3, ho, wo, ho, wo, st))),
        "warp_bilinear_motion_field": (28.0, lambda: nt.check(lib.tg_warp_bilinear(nt.ptr(hr), nt.ptr(smooth), nt.ptr(hr2), n, 3, ho, wo, ho, wo, st))),
        "warp_bilinear_random_field": (28.0, lambda: nt.check(lib.tg_warp_bilinear(nt.ptr(hr), nt.ptr(grid), nt.ptr(hr2), n, 3, ho, wo, ho, wo, st))),
        "upscale4_bilinear": (12.75, lambda: nt.check(lib.tg_upscale4_bilinear(nt.ptr(lr), nt.ptr(hr2), n, 3, h, w, 4.0, st))),
        "fused_warp_s2d_concat": (18.4, lambda: nt.check(lib.tg_fused_warp_s2d_concat(nt.ptr(lr), nt.ptr(lr_prev), nt.ptr(hr), nt.ptr(x_in), n, h, w,
                                                                                      3 * h * w, 48 * h * w, st))),
    }
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if peak else "fallback 6500 GB/s (B200_PROFILING.md)"
    peak = peak or 6500.0
    out = {"workload": f"cfg3 sizes: {n} frame(s) {w}x{h} -> {wo}x{ho}", "peak": peak, "peak_source": peak_src, "unit": "GB/s", "kernels": {}}
    for name, (bpp, fn) in ops.items():
        t = timed(fn)
        ach = bpp * hr_px / t / 1e9
        out["kernels"][name] = {"us": t * 1e6, "algorithmic_bytes_per_hr_px": bpp, "achieved": ach, "frac": ach / peak}
    return out


if __name__ == "__main__":
    print(json.dumps(measure(n=int(os.environ.get("TG_N", "2")))))
