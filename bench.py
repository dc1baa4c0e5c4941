#!/usr/bin/env python
"""bench.py — TecoGAN x4 VSR inference throughput and training-step throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): generator inference 320x180 -> 1280x720, 100-frame synthetic
clips, bf16 tensor-core convolutions, sharded by clip across GPUs (no data-path collective).
One "step" = one batch of B independent 100-frame clips per GPU through the recurrent frame loop.

  value      frames/s, whole job, LR clips already resident in HBM
  e2e        frames/s through ClipPipeline.run_host: pinned host LR in, every HR frame copied
             back to pinned host memory, copies inside the timed region
  roofline   dominant kernel (tg::frame_kernel: all 41 tcgen05 conv layers of a frame) timed per launch with CUDA events
             INSIDE the long step (one more full step right after the K timed ones, same clock / power state) against the
             sustained bf16 peak; roofline.burst = a 20-frame run timed alone against the burst peak
  train      second half of the metric ("train clips/s"): tecogan_b200.train.FRVSR_Train steps on cfg4 (N=1) and cfg5
             (global batch 32 split over the N ranks, gradient all-reduce inside the step), device-timed + e2e
  cfg3       the same measurements on BASELINE configs[2] (960x540 -> 4K, 30-frame clips): frames/s, e2e, frame-kernel
             roofline at 4K
  lr_u01     `value` re-measured with the reference-faithful U[0,1) LR pixels (93 % of the warp taps out of bounds)
  cpu_baseline / --impl reference: the CPU oracle port (torch fp32 on the host cores) on a
             bounded sample of the same workload (the SAME sample in both).
  torch_gpu_baseline: the oracle port's modules moved to this GPU and run under torch.autocast (cuDNN / ATen kernels) -
             "stock PyTorch on the same B200", the only GPU bar the reference offers (it ships no kernels).
  Only the baseline legs import oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pytorch-tecogan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

H, W, T = 180, 320, 100
FLOP_PER_LR_PIXEL = 8445312           # whole generator, MAC=2, padding not counted
FLOP_PER_LR_PIXEL_OUTCONV = 2 * 9 * 64 * 3 * 16
TRAFFIC_BYTES_PER_LAUNCH = 2.541e9     # ncu --set full, final frame_kernel, 2 clips/launch: dram read 1.337 GB + write 1.204 GB (profiles/r02_frame_final.ncu-rep)
METRIC = "720p output frames/s (x4 VSR inference)"
WORKLOAD = "cfg2: generator inference 320x180 -> 1280x720, 100-frame synthetic clips, sharded by clip"
CPU_SAMPLE_FRAMES = 3                  # bounded sample of the clip for the CPU arms (cpu_baseline AND --impl reference)
CPU_SAMPLE = f"{CPU_SAMPLE_FRAMES}-frame 320x180 clip (bounded sample of the 100-frame cfg2 clip), U[0,0.25) LR pixels"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=int(os.environ.get("TG_BENCH_CLIPS", "2")),
                    help="independent clips per GPU per step")
    ap.add_argument("--frames", type=int, default=T)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (train clips/s)")
    ap.add_argument("--no-glue", action="store_true", help="skip the HBM roofline of the glue kernels (cfg3 sizes)")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the 4K (cfg3) inference leg")
    ap.add_argument("--no-torch-gpu", action="store_true", help="skip the stock-PyTorch-on-this-GPU baseline leg")
    ap.add_argument("--e2e-format", default="u8", choices=["f32", "f16", "u8"],
                    help="what run_host ships to the host in the headline e2e (all three are reported)")
    ap.add_argument("--lib", default=None, help="measurement only: load another build of libtecogan_b200.so (same-box A/B)")
    ap.add_argument("--train-steps", type=int, default=10)
    return ap.parse_args()


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 4 + i and r[4 + i].lower().startswith("active"):
                    reasons.add(nm)
        # median over samples taken under load (upper half of the sorted clocks: idle samples at start/stop)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": (max(mx) if mx else None),
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ reference arm
def cpu_oracle_fps(n_frames, steps, warmup):
    """frames/s of the CPU oracle port (oracle/tecogan_oracle.py, torch fp32 on the host cores) on an
    n_frames-long 320x180 clip.  This is the reference's CPU path timed on this box."""
    import torch
    from oracle import synth, tecogan_oracle as O
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it may run on
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    torch.manual_seed(1)
    G = O.OracleGenerator(3, 16).eval()
    r = torch.from_numpy(synth.clip_inputs(1, n_frames, H, W, seed=1234, hi=0.25))
    for _ in range(warmup):
        O.infer_clip(G, r[:, :1])
    t0 = time.perf_counter()
    for _ in range(steps):
        O.infer_clip(G, r)
    dt = time.perf_counter() - t0
    return n_frames * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_frames = CPU_SAMPLE_FRAMES
    fps, step_s, cores = cpu_oracle_fps(n_frames, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "arm": "reference CPU path (oracle port of models.py + main.py:173-219, "
                   "torch fp32 on the host cores)",
                   "sample": CPU_SAMPLE + " per step"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": CPU_SAMPLE + f", {args.steps} steps"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_train:
        tr = {}
        for name, crop, cb in (("cfg4", 32, 4), ("cfg5", 64, 1)):
            cps, step_s, c = cpu_oracle_train(crop, cb)
            tr[name] = {"value": cps, "unit": "clips/s", "ms_per_step": step_s * 1e3, "cores": c, "kind": "port",
                        "sample": f"one oracle train step on {cb} clip(s) of 10 frames, {crop}x{crop} LR"}
        line["train"] = tr
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- training leg
# generator 8,445,312 FLOP per LR pixel forward, x3 forward+backward; discriminator 6.9018 GF per
# 128x128 sample forward, 20.20 GF forward+backward (x4 at 256x256); two discriminator passes per step.
def train_step_flops(b, t, crop):
    g = b * t * crop * crop * FLOP_PER_LR_PIXEL * 3.0
    tb = b * (3 * (t // 3)) // 3
    d = 2 * tb * 20.20e9 * (crop / 32.0) ** 2
    return g + d


def train_args(crop):
    import types
    return types.SimpleNamespace(num_resblock=16, discrim_resblocks=4, discrim_channels=128, RNN_N=10, crop_size=crop,
                                 pingpang=False, learning_rate=1e-4, vgg_scaling=-0.002, crop_dt=0.75, Dt_mergeDs=True,
                                 D_LAYERLOSS=True, EPS=1e-12, ratio=0.01, Dt_ratio_max=1.0, Dt_ratio_0=1.0,
                                 Dt_ratio_add=0.0, pp_scaling=1.0, beta=0.9, adameps=1e-8)


def cpu_oracle_train(crop, b):
    """clips/s of the CPU oracle port of the training step (oracle/train_oracle.py) — the reference's CPU path."""
    import torch
    from oracle import synth, tecogan_oracle as O, train_oracle as TO
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    torch.manual_seed(1)
    args = TO.default_train_args(crop_size=crop)
    G = O.OracleGenerator(3, 16)
    D = O.OracleDiscriminator(4, 128, 48 * (crop // 32) ** 2)
    og = torch.optim.Adam(G.parameters(), 1e-4, betas=(0.9, 0.999), eps=1e-8)
    od = torch.optim.Adam(D.parameters(), 1e-4, betas=(0.9, 0.999), eps=1e-8)
    r_in = torch.from_numpy(synth.det_uniform((b, 10, 3, crop, crop), 51, 0.0, 1.0))
    r_tg = torch.from_numpy(synth.det_uniform((b, 10, 3, 4 * crop, 4 * crop), 52, 0.0, 1.0))
    t0 = time.perf_counter()
    TO.train_step(G, D, og, od, r_in, r_tg, args, 0)
    dt = time.perf_counter() - t0
    return b / dt, dt, torch.get_num_threads()


def run_train_leg(name, crop, global_batch, world, rank, dev, steps, warmup, barrier, max_over_ranks, with_cpu, perceptual=False):
    """One training configuration: `steps` calls of tecogan_b200.train.FRVSR_Train (the reference's train entry point,
    code/train.py:374-377) on this rank's share of the global batch; gradients all-reduced inside the step when
    world > 1.  Returns the JSON sub-object (rank 0) or None."""
    import torch
    from tecogan_b200 import _native as nt, models, parallel, train as T
    lib = nt.lib()
    args = train_args(crop)
    saved_graph = T.USE_CUDA_GRAPH
    if perceptual:
        # BASELINE configs[3] wording ("random-init VGG19 perceptual loss"): the reference's VGG branch cannot run,
        # so this variant times OUR labelled, non-parity stand-in (tecogan_b200.perceptual); eager (no graph capture)
        from tecogan_b200 import perceptual as PS
        PS.ENABLED = True
        args.vgg_scaling = 0.2
        T.USE_CUDA_GRAPH = False
    per = global_batch // world
    torch.manual_seed(1)                                   # identical replicas on every rank
    G = models.generator(3, args).to(dev)
    D = models.discriminator(args).to(dev)
    parallel.broadcast_parameters(G)
    parallel.broadcast_parameters(D)
    # main.py:239-243's optimizers, built exactly as the reference builds them; tecogan_b200.train adopts them
    # (tecogan_b200.optim.FlatAdam: repo kernels on the flat buckets) and captures the step in a CUDA graph when single-process
    og = torch.optim.Adam(G.parameters(), args.learning_rate, betas=(args.beta, 0.999), eps=args.adameps)
    od = torch.optim.Adam(D.parameters(), args.learning_rate, betas=(args.beta, 0.999), eps=args.adameps)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    r_in = torch.rand((per, 10, 3, crop, crop), device=dev, generator=gen)
    r_tg = torch.rand((per, 10, 3, 4 * crop, 4 * crop), device=dev, generator=gen)
    for i in range(warmup):
        out = T.FRVSR_Train(r_in, r_tg, args, D, G, i, 0.0, 0.0, og, od)
    barrier()
    # repo kernels of the timed steps: launched directly (eager) + executed by graph replays (counted once at capture)
    l0 = lib.tg_launch_count() + T.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = T.FRVSR_Train(r_in, r_tg, args, D, G, warmup + i, 0.0, 0.0, og, od)
    e1.record()
    barrier()
    launches = lib.tg_launch_count() + T.replayed_launches - l0
    dev_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    losses = (float(out.gen_loss), float(out.d_loss))
    # end to end: pinned host batches in (H2D inside the timed region), the two losses read back every step
    h_in, h_tg = r_in.cpu().pin_memory(), r_tg.cpu().pin_memory()
    d_in, d_tg = torch.empty_like(r_in), torch.empty_like(r_tg)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        d_in.copy_(h_in, non_blocking=True)
        d_tg.copy_(h_tg, non_blocking=True)
        out = T.FRVSR_Train(d_in, d_tg, args, D, G, warmup + steps + i, 0.0, 0.0, og, od)
        host_losses = (float(out.gen_loss), float(out.d_loss))          # D2H read of the step's result
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if perceptual:
        PS.ENABLED = False
        T.USE_CUDA_GRAPH = saved_graph
    # exposed all-reduce time: the same steps with the two gradient all-reduces skipped (replicas diverge: measurement only)
    nocomm_s = None
    if world > 1:
        parallel.SKIP_ALLREDUCE = True
        T.FRVSR_Train(r_in, r_tg, args, D, G, warmup + 2 * steps, 0.0, 0.0, og, od)
        barrier()
        e0.record()
        for i in range(steps):
            T.FRVSR_Train(r_in, r_tg, args, D, G, warmup + 2 * steps + 1 + i, 0.0, 0.0, og, od)
        e1.record()
        barrier()
        nocomm_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
        parallel.SKIP_ALLREDUCE = False
    if rank != 0:
        return None
    flops = train_step_flops(global_batch, 10, crop)
    res = {"workload": name, "value": global_batch * steps / dev_s, "unit": "clips/s", "ms_per_step": dev_s / steps * 1e3,
           "global_batch": global_batch, "clips_per_gpu": per, "frames": 10, "lr_crop": crop, "steps": steps, "warmup": warmup,
           "scaling": "strong" if world > 1 or name.startswith("cfg5") else "single GPU",
           "parallelism": f"batch-data-parallel x{world}, flat-bucket gradient all-reduce (NCCL) overlapped with backward"
                          if world > 1 else "single GPU",
           "e2e": {"value": global_batch * steps / e2e_s, "unit": "clips/s",
                   "h2d_bytes_per_step": int(h_in.numel() * 4 + h_tg.numel() * 4), "d2h_bytes_per_step": 8,
                   "api": "tecogan_b200.train.FRVSR_Train (pinned host batch -> device, losses read back)"},
           "gpu_launches": int(launches),
           "step_tflops": flops / (dev_s / steps) / 1e12, "algorithmic_flops_per_step": flops,
           "losses_finite": bool(all(v == v and abs(v) != float("inf") for v in losses + host_losses)),
           "optimizer": "torch.optim.Adam objects as main.py:239-243 builds them, stepped by tecogan_b200.optim.FlatAdam (fused flat-bucket "
                        "Adam + GradScaler update + bf16 re-pack, repo kernels)" if T.FUSED_ADAM else "torch.optim.Adam + GradScaler (stock)",
           "cuda_graph": bool(T.USE_CUDA_GRAPH and T.FUSED_ADAM and (world == 1 or T.GRAPH_DATA_PARALLEL)) and not perceptual}
    if nocomm_s is not None:
        res["allreduce"] = {"bytes_per_step": int(sum(p.numel() for p in G.parameters()) + sum(p.numel() for p in D.parameters())) * 4,
                            "ms_per_step_without_allreduce": nocomm_s / steps * 1e3,
                            "exposed_ms_per_step": (dev_s - nocomm_s) / steps * 1e3,
                            "how": "the same steps timed again with both gradient all-reduces skipped (tecogan_b200.parallel.SKIP_ALLREDUCE)"}
    if perceptual:
        res["perceptual_loss"] = ("NON-PARITY stand-in (tecogan_b200.perceptual): random-init VGG19-to-conv4_4 on the conv core, features "
                                  "conv2_2/3_4/4_4, cosine loss, vgg_scaling 0.2; the reference's VGG branch is unrunnable (SURVEY.md 8c). "
                                  "step_tflops counts the generator + discriminator work only")
    if with_cpu:
        cb = 4 if crop == 32 else 1
        cps, step_s, cores = cpu_oracle_train(crop, cb)
        res["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port",
                               "sample": f"one oracle train step (oracle/train_oracle.py, torch CPU fp32) on {cb} clip(s) of "
                                         f"10 frames, {crop}x{crop} LR"}
    return res


# ------------------------------------------------------------------------------------- ours
def torch_gpu_fps(dev, n_frames, h, w, steps=3, warmup=1):
    """"Stock PyTorch on the same B200": the oracle port's generator (plain nn.Conv2d / nn.ConvTranspose2d modules, the
    reference's own layer stack) on this GPU under torch.autocast(fp16) - cuDNN / ATen kernels, exactly the arithmetic
    the reference's GPU path runs (main.py:171-172) - through the reference's frame loop (main.py:173-219: F.grid_sample
    with the .half() grid, deprocess, space-to-depth, cat).  Kinder to the baseline than the reference is to itself:
    everything stays on the device (the reference bounces every frame through host memory, main.py:195,203,214)."""
    import torch
    import torch.nn.functional as F
    from oracle import tecogan_oracle as O
    torch.manual_seed(1)
    G = O.OracleGenerator(3, 16).to(dev).eval()
    gen = torch.Generator(device=dev).manual_seed(1234)
    r = torch.rand((1, n_frames, 3, h, w), device=dev, generator=gen) * 0.25

    @torch.no_grad()
    def clip():
        with torch.autocast("cuda", dtype=torch.float16):
            flow = F.interpolate(r[0, :-1] * 4.0, scale_factor=4, mode="bilinear", align_corners=False)[:, 0:2]   # main.py:186-189
            x = torch.cat((r[:, 0], torch.zeros((1, 48, h, w), device=dev)), dim=1)                                # :191-193
            prev = G(x).view(1, 3, 4 * h, 4 * w)                                                                   # :195-196
            outs = [prev]
            for i in range(n_frames - 1):                                                                          # :199
                grid = flow[i].contiguous().view(1, 4 * h, 4 * w, 2)                                               # :200-201
                wp = F.grid_sample(prev.float(), grid.half().float(), mode="bilinear", padding_mode="zeros", align_corners=False)  # :203
                wp = (wp + 1) / 2                                                                                  # :206
                x = torch.cat((r[:, i + 1], F.pixel_unshuffle(wp, 4)), dim=1)                                      # :207-213
                prev = G(x)                                                                                        # :214
                outs.append(prev)
            return torch.stack(outs, dim=1)

    for _ in range(warmup):
        clip()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = clip()
    e1.record()
    torch.cuda.synchronize()
    s = e0.elapsed_time(e1) * 1e-3
    del G, out
    torch.cuda.empty_cache()
    return n_frames * steps / s, s / steps


def inference_leg(G, dev, world, rank, B, frames, h, w, K, Wm, lr_hi, barrier, max_over_ranks, do_e2e, e2e_format, do_roof,
                  traffic_per_launch, sampler=None):
    """One inference configuration: B independent `frames`-frame h x w clips per GPU through the recurrent loop.
    Returns (value, ms_per_step, launches, e2e dict, roofline dict, finite, clocks)."""
    import ctypes
    import torch
    from tecogan_b200 import _native as nt
    from tecogan_b200.pipeline import ClipPipeline
    lib = nt.lib()
    pipe = ClipPipeline(G, B, frames, h, w, dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    lr = torch.rand((B, frames, 3, h, w), device=dev, generator=gen) * lr_hi
    out = torch.empty((B, frames, 3, 4 * h, 4 * w), dtype=torch.float32, device=dev)
    for _ in range(Wm):
        pipe.run_device(lr, out)
    barrier()
    if sampler:
        sampler.start()
    launches0 = lib.tg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        pipe.run_device(lr, out)
    e1.record()
    barrier()
    launches = lib.tg_launch_count() - launches0
    dev_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    def profile_frames(pp, lr_in, nframes, peak_key, peak_fallback, peak_src, out_buf=None):
        """Per-launch CUDA events (library hooks tg_profile_begin/end) over one run of `nframes` frames -> roofline dict."""
        nt.check(lib.tg_profile_begin())
        pp.run_device(lr_in, out_buf)
        cap = 64 * nframes + 64
        ids = (ctypes.c_int * cap)()
        ms = (ctypes.c_float * cap)()
        work = (ctypes.c_double * cap)()
        n = lib.tg_profile_end(cap, ids, ms, work)
        all_ms = sum(ms[i] for i in range(n))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, src = peaks.get(peak_key), peak_src
        if not peak:
            peak, src = peak_fallback, "fallback (B200_PROFILING.md)"
        fr = [i for i in range(n) if ids[i] == 5]
        if fr:      # frame kernel: one launch = all 41 conv layers of one generator forward for B clips
            k_ms = sum(ms[i] for i in fr)
            k_n = len(fr)
            flops = float(k_n) * B * FLOP_PER_LR_PIXEL * h * w
            kname = "tg::frame_kernel (persistent tcgen05 implicit-GEMM generator forward, 1 launch/frame)"
        else:       # per-layer path: 40 conv_tc_kernel<64> launches per frame
            k_ms = sum(ms[i] for i in range(n) if ids[i] == 0)
            k_n = sum(1 for i in range(n) if ids[i] == 0)
            flops = float(nframes) * B * (FLOP_PER_LR_PIXEL - FLOP_PER_LR_PIXEL_OUTCONV) * h * w
            kname = "tg::conv_tc_kernel<64> (tcgen05 implicit-GEMM conv, 40 launches/frame)"
        achieved = flops / (k_ms * 1e-3) / 1e12
        return {"kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "peak_source": src,
                "traffic": traffic_per_launch if fr else None,
                "traffic_source": ("ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one frame_kernel launch at this "
                                   "configuration (profiles/)") if (fr and traffic_per_launch) else None,
                "launches_timed": k_n, "avg_launch_us": k_ms * 1e3 / max(k_n, 1),
                "algorithmic_flops_per_launch": flops / max(k_n, 1),
                "kernel_share_of_step": k_ms / all_ms if all_ms else None}

    # roofline of the dominant kernel INSIDE the long step: one more full step, immediately after the K timed ones (same clock /
    # power state), with a CUDA-event pair around every launch; peak = the sustained figure
    roof = None
    if do_roof and rank == 0:
        roof = profile_frames(pipe, lr, frames, "bf16_tflops_sustained", 1400.0,
                              "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step: one full step right "
                              "after the timed ones)", out)
    clocks = sampler.stop() if sampler else None
    value = world * B * frames * K / dev_s
    finite = bool(torch.isfinite(out[:, -1]).all().item())
    del out

    # ---------------- end to end with host buffers ----------------
    e2e = None
    if do_e2e:
        dt = {"f32": torch.float32, "f16": torch.float16, "u8": torch.uint8}
        lr_host = torch.empty((B, frames, 3, h, w), dtype=torch.float32).pin_memory()
        lr_host.copy_(lr.cpu())
        res = {}
        for fmt in ("f32", "f16", "u8"):
            out_host = torch.empty(pipe._out_shape(dt[fmt]), dtype=dt[fmt]).pin_memory()
            for _ in range(max(1, Wm - 1)):
                pipe.run_host(lr_host, out_host, out_dtype=dt[fmt])
            barrier()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(K):
                pipe.run_host(lr_host, out_host, out_dtype=dt[fmt])           # returns with every frame in host memory
            e1.record()
            barrier()
            host_s = max_over_ranks(max(e0.elapsed_time(e1) * 1e-3, 0.0))
            wall_s = max_over_ranks(time.perf_counter() - t0)
            bi, bo = pipe.bytes_per_run(dt[fmt])
            res[fmt] = {"value": world * B * frames * K / max(host_s, wall_s), "unit": "frames/s", "h2d_bytes_per_step": bi,
                        "d2h_bytes_per_step": bo, "result_mean": float(out_host[-1].float().mean()) / (255.0 if fmt == "u8" else 1.0)}
            del out_host
        what = {"f32": "f32 planar frames (the parity default, bit-identical to generator.infer_clip)",
                "f16": "fp16 planar frames (the f32 result rounded to fp16: what the reference's autocast GPU path emits, main.py:171-172)",
                "u8": "uint8 NHWC frames ((x*255) truncated: exactly what the reference's save_as_gif makes of the clip before "
                      "writing it, code/ops.py:234-237)"}
        e2e = dict(res[e2e_format])
        e2e["api"] = ("tecogan_b200.pipeline.ClipPipeline.run_host (pinned host LR in, every HR frame copied back to pinned host "
                      "memory on a side stream while the next frame computes; copies inside the timed region)")
        e2e["output_format"] = what[e2e_format]
        e2e["other_formats"] = {k: {"value": v["value"], "d2h_bytes_per_step": v["d2h_bytes_per_step"], "output_format": what[k]}
                                for k, v in res.items() if k != e2e_format}
        del lr_host

    # ---------------- roofline of the dominant kernel: a short run timed alone, against the BURST peak ----------------
    if roof is not None:
        pf = min(frames, 20)
        pipe_p = ClipPipeline(G, B, pf, h, w, dev) if pf != frames else pipe
        lr_p = lr[:, :pf].contiguous()
        pipe_p.run_device(lr_p)
        torch.cuda.synchronize()
        burst = profile_frames(pipe_p, lr_p, pf, "bf16_tflops", 1650.0,
                               "MEASURED_PEAKS.json bf16_tflops (burst: a 20-frame run timed alone after an idle gap)")
        roof["burst"] = {k: burst[k] for k in ("achieved", "peak", "frac", "peak_source", "launches_timed", "avg_launch_us")}
    del pipe, lr
    torch.cuda.empty_cache()
    return value, dev_s / K * 1e3, int(launches), e2e, roof, finite, clocks


def run_ours(args):
    import types
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from tecogan_b200 import _native as nt, models
    if args.lib:                                           # measurement only (same-box A/B of two builds)
        nt.LIB_PATH = os.path.abspath(args.lib)
    nt.lib()

    B, K, Wm = args.clips, args.steps, max(args.warmup, 3)
    frames = args.frames
    torch.manual_seed(1)                                   # reference default --rand_seed 1 (main.py:34)
    G = models.generator(3, types.SimpleNamespace(num_resblock=16)).to(dev).eval()   # random init

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- cfg2: the headline (value, e2e, roofline) ----------------
    value, ms_step, launches, e2e, roof, finite, clocks = inference_leg(
        G, dev, world, rank, B, frames, H, W, K, Wm, 0.25, barrier, max_over_ranks, not args.no_e2e, args.e2e_format, True,
        TRAFFIC_BYTES_PER_LAUNCH if B == 2 else None, ClockSampler(local))

    # ---------------- cfg2 with the reference-faithful LR distribution U[0,1) (value only) ----------------
    v_u01, ms_u01, _, _, _, fin_u01, _ = inference_leg(G, dev, world, rank, B, min(frames, 30), H, W, max(2, K // 2), 3, 1.0, barrier,
                                                       max_over_ranks, False, args.e2e_format, False, None)
    lr_u01 = {"value": v_u01, "unit": "frames/s", "lr_dist": "U[0,1): the reference-faithful input (4*LR as the 'flow' puts 93 % of "
              "the warp taps out of bounds -> zeros)", "frames_per_clip": min(frames, 30), "outputs_finite": fin_u01}

    # ---------------- cfg3: 960x540 -> 4K, 30-frame clips ----------------
    cfg3 = None
    if not args.no_cfg3:
        v3, ms3, l3, e3, r3, f3, _ = inference_leg(G, dev, world, rank, 1, 30, 540, 960, max(2, K // 2), 3, 0.25, barrier, max_over_ranks,
                                                   not args.no_e2e, args.e2e_format, True, None)
        cfg3 = {"workload": "cfg3: generator inference 960x540 -> 3840x2160 (4K), 30-frame synthetic clips, 1 clip per GPU per step",
                "metric": "4K output frames/s (x4 VSR inference)", "value": v3, "unit": "frames/s", "ms_per_step": ms3,
                "720p_equivalent_frames_per_s": v3 * 9.0, "gpu_launches": l3, "e2e": e3, "roofline": r3, "outputs_finite": f3,
                "lr_dist": "U[0,0.25)"}

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, step_s, cores = cpu_oracle_fps(CPU_SAMPLE_FRAMES, 1, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": CPU_SAMPLE + " through oracle.infer_clip (torch CPU fp32); the same sample as --impl reference"}

    # ---------------- stock PyTorch (cuDNN, autocast fp16) on this GPU: the bar the reference's own GPU path sets ----------------
    tgpu = None
    if rank == 0 and world == 1 and not args.no_torch_gpu:
        try:
            fps2, step2 = torch_gpu_fps(dev, 10, H, W)
            tgpu = {"value": fps2, "unit": "frames/s", "kind": "oracle-port modules on cuda:0 under torch.autocast(fp16), device-resident "
                    "loop (no PCIe hops)", "sample": "one 10-frame 320x180 clip per step, 3 steps", "ms_per_frame": step2 / 10 * 1e3,
                    "ours_over_this": value / world / fps2}
        except Exception as e:                               # a baseline leg must never take the bench down
            tgpu = {"unavailable": repr(e)[:200]}

    # ---------------- HBM roofline of the memory-bound glue kernels at cfg3 (4K) sizes ----------------
    glue = None
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_glue:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import glue_bench
        glue = glue_bench.measure(n=2, dev=dev)
        torch.cuda.empty_cache()

    # ---------------- training step (BASELINE.json metric, second half: train clips/s) ----------------
    train = None
    if not args.no_train:
        train = {}
        with_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
        if world == 1:
            train["cfg4"] = run_train_leg("cfg4: training step, batch 4 x 10 frames of 32x32 LR crops, single B200", 32, 4,
                                          world, rank, dev, args.train_steps, 3, barrier, max_over_ranks, with_cpu)
            train["cfg4_perceptual_standin"] = run_train_leg(
                "cfg4 + perceptual-loss stand-in (non-parity): batch 4 x 10 frames of 32x32 LR crops, single B200", 32, 4, world, rank, dev,
                max(3, args.train_steps // 2), 3, barrier, max_over_ranks, False, perceptual=True)
        train["cfg5"] = run_train_leg("cfg5: data-parallel training, global batch 32 x 10 frames of 64x64 LR crops", 64, 32,
                                      world, rank, dev, max(3, args.train_steps // 2), 3, barrier, max_over_ranks, with_cpu)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "clips_per_gpu_per_step": B, "frames_per_clip": frames, "lr_dist": "U[0,0.25) (all warp taps "
                       "in bounds)", "weights": "random init, seed 1",
                       "generator_mode": {0: "one launch per conv layer", 2: "persistent frame kernel"}.get(int(G.amode), str(G.amode)),
                       "parallelism": f"clip-sharded x{world}, no collective",
                       "l2": "no explicit flush: every frame streams ~0.56 GB of activations per clip through the "
                             "126 MB L2, far larger than L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "torch_gpu_baseline": tgpu, "outputs_finite": finite, "lr_u01": lr_u01, "cfg3": cfg3, "glue": glue, "train": train,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
