"""bench.py -- denoised frames/sec of the MuDG sampler hot path on B200 (BASELINE.json metric).

One "step" = one clip per rank: DDIMSampler.sample(S=50, CFG 7.5, guidance_rescale 0.7, eta 1, uniform_trailing) on a
[1,4,16,72,128] latent (MDM1024, BASELINE configs[2]/[3]) + decode_first_stage -> 16 frames of 576x1024.  Weights are
seeded-random of the reference architecture (no checkpoint is reachable), conditioning is synthetic.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config mdm1024|mdm512|mdm1024_t64]
                  [--no-extras] [--no-cpu-baseline]

Prints ONE JSON line (rank 0).
  value     whole-job frames/s with every clip's conditioning already resident in HBM (CUDA-graph replay of the forward);
            max over ranks of the CUDA-event time.
  e2e       BASELINE config 4 as the driver runs it (virtual_pose_render.py:187-274): rank 0 holds the item list (one item
            per clip: steps x ranks clips), `shard.scatter_items` hands every rank its share, each clip's conditioning
            is copied H2D from pinned host memory, sampled, decoded, converted to uint8 frames on the GPU
            (mudg_postdecode), gathered to rank 0 over NCCL (`shard.gather_frames_to`) and read back D2H there -- all
            inside the timed region.
  roofline  top level = the whole path (algorithmic FLOPs of the reference graph / wall time against the measured
            sustained dense peak); `kernels` = one entry per kernel family, timed live with CUDA events around every
            launch of one extra (eager) clip, each against the roofline that bounds it (tensor or HBM).
  extra     (N = 1 only) driver-run numbers for BASELINE configs 2 (MDM512) and 5 (MDM1024, T = 64).
--impl reference: the UNCHANGED reference modules (baseline/_ref, copied from /root/reference by __graft_entry__.build())
timed on the host cores: one full-resolution UNetModel.forward + one AutoencoderKL.decode frame of the SAME config, scaled
by the number of forwards / frames per clip (no extrapolation across resolutions); rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (yaml, T, h, w, steps, cfg_scale, guidance_rescale, UNet TFLOP per forward (B=1), VAE TFLOP per frame)
    "mdm1024": ("stage2-1024_mdm_waymo_infer_synthetic.yaml", 16, 72, 128, 50, 7.5, 0.7, 52.340, 5.7543),
    "mdm512": ("stage1-512_mdm_waymo_infer_synthetic.yaml", 16, 40, 64, 50, 1.0, 0.0, 12.604, 1.5635),
    "mdm1024_t64": ("stage2-1024_mdm_waymo_infer_synthetic.yaml", 64, 72, 128, 50, 7.5, 0.7, 212.49, 5.7543),
}
METRIC = "denoised frames/sec @576x1024x16f, 50 DDIM steps"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1400.0, 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.25)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ reference (CPU) legs
def _cpu_reference_models(unet_sd, vae_sd):
    """The unchanged reference UNetModel / AutoencoderKL on the CPU (fp32), or -- when baseline/_ref is absent -- the oracle
    port.  Returns (kind, unet_fn(x, ts, lab, ctx, fs), vae_fn(z))."""
    import torch
    from oracle import mudg_oracle as O
    from oracle import refimpl
    import contextlib
    cfg, vcfg = O.UNetCfg(), O.VaeCfg()
    if refimpl.ref_root() is not None:
        with contextlib.redirect_stdout(sys.stderr):       # the reference prints while constructing; stdout carries ONE JSON line
            m = refimpl.reference_unet(cfg, None, device="meta")
            m.load_state_dict(unet_sd, strict=True, assign=True)
            vae_fn = None
            if vae_sd is not None:
                v = refimpl.reference_vae(vcfg, None, device="meta")
                v.load_state_dict(vae_sd, strict=True, assign=True)
                vae_fn = v.decode
        unet_fn = lambda x, ts, lab, ctx, fs: m(x, ts, c_label=lab, context=ctx, fs=fs)      # noqa: E731
        return "reference", unet_fn, vae_fn
    unet_fn = lambda x, ts, lab, ctx, fs: O.unet_forward(unet_sd, cfg, x, ts, lab, ctx, fs)   # noqa: E731
    vae_fn = (lambda z: O.vae_decode(vae_sd, vcfg, z)) if vae_sd is not None else None
    return "port", unet_fn, vae_fn


def _cpu_inputs(T, h, w):
    import torch
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 12, T, h, w, generator=g)
    ctx = torch.randn(1, 77 + (16 * T if T == 16 else 256), 1024, generator=g)
    return x, torch.full((1,), 500, dtype=torch.long), torch.zeros(1, dtype=torch.long), ctx, torch.full((1,), 10, dtype=torch.long)


def _seeded_cpu_weights(vae=True):
    from oracle import mudg_oracle as O
    sd = O.seeded_state_dict(O.unet_param_shapes(O.UNetCfg()), seed=0)
    vsd = O.seeded_state_dict(O.vae_param_shapes(O.VaeCfg()), seed=1) if vae else None
    return sd, vsd


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path, all host threads, on OUR config."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    _, T, h, w, S, cfg_scale, _, unet_tf, vae_tf = CONFIGS[args.config]
    evals = S * (2 if cfg_scale != 1.0 else 1)
    sd, vsd = _seeded_cpu_weights()
    kind, unet_fn, vae_fn = _cpu_reference_models(sd, vsd)
    # the reference's einsum attention materialises (frames x heads) x HW^2 fp32 scores and their softmax: bound the memory
    frames, hw = T, h * w
    need_gb = 2 * frames * 5 * hw * hw * 4 / 1e9 + 30
    try:
        import psutil
        avail_gb = psutil.virtual_memory().available / 1e9
    except Exception:
        avail_gb = 0.0
    same_config = kind == "port" or avail_gb > need_gb
    sT, sh, sw = (T, h, w) if same_config else (16, 40, 64)
    with torch.no_grad():
        unet_fn(*_cpu_inputs(4, 8, 8))                                   # page in weights / thread pool (not a sample)
        t0 = time.time()
        unet_fn(*_cpu_inputs(sT, sh, sw))
        t_fwd = time.time() - t0
        t0 = time.time()
        vae_fn(torch.randn(1, 4, sh, sw))
        t_dec = time.time() - t0
    scale_u = 1.0 if same_config else unet_tf / CONFIGS["mdm512"][7] * (T / 16)
    scale_v = 1.0 if same_config else vae_tf / CONFIGS["mdm512"][8]
    clip_s = evals * t_fwd * scale_u + T * t_dec * scale_v
    v = T / clip_s
    what = "UNCHANGED reference modules (baseline/_ref: lvdm UNetModel + AutoencoderKL)" if kind == "reference" else "oracle port (baseline/_ref absent)"
    sample = (f"{what}, CPU fp32, {threads} threads: ONE UNetModel.forward on [1,12,{sT},{sh},{sw}] = {t_fwd:.1f} s and ONE "
              f"AutoencoderKL.decode frame at {8 * sh}x{8 * sw} = {t_dec:.1f} s; clip = {evals} forwards + {T} frames"
              + ("" if same_config else f"; host RAM {avail_gb:.0f} GB < {need_gb:.0f} GB needed by the reference's einsum attention at "
                 f"{h}x{w}: sampled at MDM512 size and scaled by the FLOP ratio"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
        "ms_per_step": clip_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": args.config, "same_config": same_config,
                                        "note": "reference path on the host cores; a bounded sample of the clip, see cpu_baseline.sample"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_leg(model, cfgname):
    """cpu_baseline of our arm: ONE reference UNetModel.forward at MDM512 size (~10-30 s of CPU work) on the weights of the
    benchmarked model, scaled to the clip by forwards per clip and the FLOP ratio of the two latent sizes."""
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    _, T, h, w, S, cfg_scale, _, unet_tf, vae_tf = CONFIGS[cfgname]
    evals = S * (2 if cfg_scale != 1.0 else 1)
    sd = {k: v.detach().float().cpu() for k, v in model.model.diffusion_model.state_dict().items()}
    kind, unet_fn, _ = _cpu_reference_models(sd, None)
    with torch.no_grad():
        unet_fn(*_cpu_inputs(4, 8, 8))
        t0 = time.time()
        unet_fn(*_cpu_inputs(16, 40, 64))
        t_fwd = time.time() - t0
    scale = unet_tf / CONFIGS["mdm512"][7]
    clip_s = evals * t_fwd * scale * (1.0 + T * vae_tf / (evals * unet_tf))      # decode share by FLOPs
    return {"value": T / clip_s, "unit": "frames/s", "cores": threads, "kind": kind,
            "sample": f"one {'unchanged reference' if kind == 'reference' else 'oracle-port'} UNetModel.forward, CPU fp32, latent [1,12,16,40,64]: "
                      f"{t_fwd:.1f} s; x{scale:.2f} (FLOP ratio to this config's latent) x{evals} forwards per clip, decode added by its FLOP share"}


# ------------------------------------------------------------------------------------------------ our arm
def build_model(cfgname, device, share=None):
    import torch
    from mudg_b200 import compat
    compat.install()
    from omegaconf import OmegaConf
    from utils.utils import instantiate_from_config
    yaml_name = CONFIGS[cfgname][0]
    cfg = OmegaConf.load(os.path.join(ROOT, "configs", yaml_name)).model
    cfg.params.unet_config.params.use_checkpoint = False
    torch.manual_seed(0)
    model = instantiate_from_config(cfg)
    if share is not None:                  # same architecture and weights (the two infer yamls differ in the schedule only)
        model.model.diffusion_model = share.model.diffusion_model
        model.first_stage_model = share.first_stage_model
    else:
        # seeded, nowhere-zero weights (the reference zero-initialises several layers: SURVEY.md App. D #1)
        g = torch.Generator().manual_seed(0)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if p.dim() == 1:
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g) if name.endswith("weight") else 0.02 * torch.randn(p.shape, generator=g))
                elif float(p.abs().sum()) == 0.0:
                    fan = p[0].numel()
                    p.copy_(torch.randn(p.shape, generator=g) / fan ** 0.5)
    model = model.to(device).eval()
    model.perframe_ae = True
    return model


class Runner:
    """One config on one rank: conditioning factory + the clip function (the call a user makes)."""

    def __init__(self, model, cfgname, dev):
        import torch
        from lvdm.models.samplers.ddim import DDIMSampler
        self.torch, self.model, self.dev = torch, model, dev
        _, self.T, self.h, self.w, self.S, self.cfg_scale, self.g_rescale, self.unet_tf, self.vae_tf = CONFIGS[cfgname]
        self.sampler = DDIMSampler(model)
        self.B = 1
        self.label = torch.zeros(self.B, 1, dtype=torch.long, device=dev)
        self.fs = torch.full((self.B,), 10, dtype=torch.long, device=dev)

    def make_cond(self, seed, pin=False):
        torch = self.torch
        g = torch.Generator().manual_seed(seed)
        ctx_len = 77 + 16 * self.T if self.T == 16 else 77 + 256
        t = dict(ctx=torch.randn(self.B, ctx_len, 1024, generator=g), uc=torch.randn(self.B, ctx_len, 1024, generator=g),
                 cat=0.5 * torch.randn(self.B, 8, self.T, self.h, self.w, generator=g))
        if pin:
            return {k: v.pin_memory() for k, v in t.items()}
        return {k: v.to(self.dev) for k, v in t.items()}

    def clip(self, c, seed, steps=None):
        torch = self.torch
        torch.manual_seed(seed)
        cond = {"c_crossattn": [c["ctx"]], "c_concat": [c["cat"]]}
        uc = {"c_crossattn": [c["uc"]], "c_concat": [c["cat"]]} if self.cfg_scale != 1.0 else None
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            z, _ = self.sampler.sample(S=steps or self.S, conditioning=cond, batch_size=self.B, shape=[4, self.T, self.h, self.w],
                                       verbose=False, unconditional_guidance_scale=self.cfg_scale, unconditional_conditioning=uc,
                                       eta=1.0, cfg_img=None, mask=None, x0=None, fs=self.fs, timestep_spacing="uniform_trailing",
                                       guidance_rescale=self.g_rescale, sparse_x=None, class_label=self.label,
                                       unconditional_conditioning_img_nonetext=None)
            return self.model.decode_first_stage(z)

    def flops_clip(self):
        evals = self.S * (2 if self.cfg_scale != 1.0 else 1)
        return (evals * self.unet_tf * self.B + self.B * self.T * self.vae_tf) * 1e12


def profile_report(L):
    n = L.mudg_profile_report(None, 0)
    buf = ctypes.create_string_buffer(int(n) + 16)
    L.mudg_profile_report(buf, ctypes.c_size_t(int(n) + 16))
    rows = []
    for line in buf.value.decode().strip().splitlines()[1:]:
        fam, shape, launches, ms, flops, nbytes = line.split(",")
        rows.append(dict(family=fam, shape=shape, launches=int(float(launches)), ms=float(ms), flops=float(flops), bytes=float(nbytes)))
    return rows


def kernel_table(rows, step_ms, sus, hbm):
    """Per kernel family: time share of the instrumented step and achieved rate against the roofline that bounds it.
    A GEMM shape whose arithmetic intensity is below the machine balance (sustained FLOP/s / HBM B/s) is HBM-bound."""
    balance = sus * 1e12 / (hbm * 1e9)
    fams = {}
    for r in rows:
        bound = "tensor" if r["flops"] > 0 and (r["bytes"] <= 0 or r["flops"] / r["bytes"] >= balance) else "hbm"
        key = r["family"] if r["family"] != "gemm" else f"gemm ({bound}-bound shapes)"
        f = fams.setdefault(key, dict(family=key, bound=bound, launches=0, ms=0.0, flops=0.0, bytes=0.0))
        f["launches"] += r["launches"]; f["ms"] += r["ms"]; f["flops"] += r["flops"]; f["bytes"] += r["bytes"]
    out = []
    for f in sorted(fams.values(), key=lambda f: -f["ms"]):
        if f["ms"] <= 0:
            continue
        e = {"kernel": f["family"], "bound": f["bound"], "launches": f["launches"], "ms_per_step": f["ms"],
             "share_of_step": f["ms"] / step_ms}
        if f["bound"] == "tensor":
            a = f["flops"] / f["ms"] / 1e9
            e.update(achieved=a, peak=sus, unit="TFLOP/s", frac=a / sus)
        elif f["bytes"] > 0:
            a = f["bytes"] / f["ms"] / 1e6
            e.update(achieved=a, peak=hbm, unit="GB/s", frac=a / hbm)
        out.append(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="mdm1024", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--dump-shapes", default=None, help="write the per-(family, shape) profile rows of the instrumented clip to this CSV")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from mudg_b200 import shard
    from mudg_b200._lib import lib
    from mudg_b200.engine import postdecode, MUDG_POST_COLOR

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(args.config, dev)
    R = Runner(model, args.config, dev)
    K, W = max(args.steps, 1), max(args.warmup, 0)
    T, h, w = R.T, R.h, R.w

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(ms):
        return shard.max_over_ranks(ms, device=dev)

    # ---- the item list (BASELINE config 4): steps x ranks clips, clip i seeded 123 + i; rank 0 owns the list
    items = [{"id": i, "traj": f"clip{i:04d}", "seed": 123 + i} for i in range(K * world)] if rank == 0 else None
    mine = shard.scatter_items(items, key=lambda it: it["traj"])
    assert len(mine) == K, (len(mine), K)

    warm = R.make_cond(99 + rank)
    for i in range(W):
        R.clip(warm, 99 + rank)
    eng = model.model.diffusion_model.engine()
    veng = model.first_stage_model.engine()

    # ---- timed: every clip's conditioning resident in HBM (CUDA-graph replay of the UNet forward); each clip has its own
    # conditioning, so the cross-attention K/V projection (mudg_set_context) runs once per clip inside the region
    dev_conds = [R.make_cond(it["seed"]) for it in mine]
    L = lib()
    L.mudg_profile_report.restype = ctypes.c_size_t
    clock = ClockSampler(local)
    clock.start()
    l0 = eng.launch_count() + veng.launch_count()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it, c in zip(mine, dev_conds):
        frames = R.clip(c, it["seed"])
    e1.record()
    sync_all()
    ms = reduce_max(e0.elapsed_time(e1))
    clocks = clock.stop()
    launches = eng.launch_count() + veng.launch_count() - l0
    del dev_conds

    # ---- e2e (config 4 as the driver runs it): scatter the item list, per clip: pinned host conditioning -> H2D, sample,
    # decode, uint8 frames on the GPU, gather to rank 0 over NCCL, D2H on rank 0
    K2 = min(K, 5)
    host_conds = [R.make_cond(it["seed"], pin=True) for it in mine[:K2]]
    out_host = torch.empty((world, R.B, T, 3, 8 * h, 8 * w), dtype=torch.uint8).pin_memory() if rank == 0 else None
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gather_ms = 0.0
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    items2 = [{"id": i, "traj": f"clip{i:04d}", "seed": 123 + i} for i in range(K2 * world)] if rank == 0 else None
    mine2 = shard.scatter_items(items2, key=lambda it: it["traj"])
    for it, hc in zip(mine2, host_conds):
        c = {k: v.to(dev, non_blocking=True) for k, v in hc.items()}
        fr = R.clip(c, it["seed"])
        rgb, _, _ = postdecode(fr, [MUDG_POST_COLOR] * R.B)
        g0.record()
        allf, ids = shard.gather_frames_to(rgb[None], torch.tensor([it["id"]], device=dev), dst=0)
        g1.record()
        if rank == 0:
            out_host.copy_(allf, non_blocking=True)
        g1.synchronize()
        gather_ms += g0.elapsed_time(g1)
    e3.record()
    sync_all()
    ms_e2e = reduce_max(e2.elapsed_time(e3))
    h2d = sum(v.numel() * v.element_size() for v in host_conds[0].values())
    d2h = world * R.B * T * 3 * 8 * h * 8 * w            # rank 0 reads every rank's uint8 frames of the step
    gather_bytes = d2h if world > 1 else 0

    # ---- roofline leg: ONE extra clip with a CUDA-event pair around every launch of every kernel family (eager forwards,
    # so this clip is not part of `value`)
    L.mudg_profile(1)
    sync_all()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    R.clip(warm, 99 + rank)
    e5.record()
    sync_all()
    ms_prof = e4.elapsed_time(e5)
    rows = profile_report(L)
    L.mudg_profile(0)
    if args.dump_shapes and rank == 0:
        with open(args.dump_shapes, "w") as f:
            f.write("family,shape,launches,ms,flops,bytes,tflops,gbs\n")
            for r in sorted(rows, key=lambda r: -r["ms"]):
                f.write(f"{r['family']},{r['shape']},{r['launches']},{r['ms']:.4f},{r['flops']:.4e},{r['bytes']:.4e},"
                        f"{r['flops'] / max(r['ms'], 1e-9) / 1e9:.1f},{r['bytes'] / max(r['ms'], 1e-9) / 1e6:.1f}\n")

    frames_total = world * K * R.B * T
    value = frames_total / (ms / 1e3)
    e2e_value = world * K2 * R.B * T / (ms_e2e / 1e3)

    # ---- extras (N = 1): BASELINE configs 2 and 5 through the same call, driver-run
    extra = {}
    if world == 1 and not args.no_extras and args.config == "mdm1024":
        sus0 = peaks()[0]
        for name, nclips in (("mdm512", 2), ("mdm1024_t64", 1)):
            try:
                m2 = model if name == "mdm1024_t64" else build_model(name, dev, share=model)
                R2 = Runner(m2, name, dev)
                c2 = R2.make_cond(7)
                R2.clip(c2, 7, steps=4)                                  # eager / capture / replay of the new shape
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(nclips):
                    R2.clip(c2, 7 + i)
                b.record()
                torch.cuda.synchronize()
                ms2 = a.elapsed_time(b)
                tf2 = R2.flops_clip() * nclips / (ms2 / 1e3) / 1e12
                extra[name] = {"value": nclips * R2.T / (ms2 / 1e3), "unit": "frames/s", "clips": nclips, "ms_per_clip": ms2 / nclips,
                               "workload": f"latent [1,4,{R2.T},{R2.h},{R2.w}], {R2.S} DDIM steps, CFG {R2.cfg_scale}, + VAE decode of {R2.T} frames",
                               "path_tflops": tf2, "path_frac_of_sustained_peak": tf2 / sus0}
            except Exception as e:                                        # an extra must never cost the headline
                extra[name] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()

    if rank == 0:
        sus, burst, hbm, how = peaks()
        flops_clip = R.flops_clip()
        path_tf = flops_clip * K / (ms / 1e3) / 1e12
        kernels = kernel_table(rows, ms_prof, sus, hbm)
        try:      # ncu --set full captures of this round (dram__bytes_read + write per launch), committed under profiles/
            with open(os.path.join(ROOT, "profiles", "r2_ncu_kernels.json")) as f:
                ncu = json.load(f)
        except Exception:
            ncu = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_parity_full.json")) as f:
                pf = json.load(f)
            parity = {"source": "profiles/r2_parity_full.json (tests/gpu_parity_full.py on B200, fp32 truth = unchanged reference / oracle)",
                      "unet_mdm512": pf.get("unet40", {}).get("ours_vs_truth"), "unet_mdm1024_cfg": pf.get("unet72", {}).get("ours_vs_truth"),
                      "reference_fp16_gap_mdm512": pf.get("unet40", {}).get("reference_fp16_vs_truth"),
                      "clip_50_steps_mdm512_latent": pf.get("sample512", {}).get("ours_latent_vs_truth"),
                      "clip_50_steps_mdm512_frames": pf.get("sample512", {}).get("ours_frames_vs_truth")}
        except Exception:
            parity = None
        res = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 (fp32 accumulate; norms/softmax fp32)", "data": "synthetic",
            "config": {"workload": f"{args.config}: latent [1,4,{T},{h},{w}], {R.S} DDIM steps, CFG {R.cfg_scale} "
                                   f"(cond+uncond batched as N=2, shared prefix), guidance_rescale {R.g_rescale}, eta 1.0, + VAE decode of {T} frames; "
                                   f"item list of {K * world} clips sharded over {world} rank(s) (BASELINE config 4 at N > 1)",
                       "clips_per_rank_per_step": 1, "weights": "seeded random, reference architecture (1.44 B param UNet)",
                       "l2": "working set per UNet forward (9.5 GB) >> 126 MB L2", "timing": "CUDA events, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": K2,
                    "gather_ms_per_step": gather_ms / K2, "gather_bytes_per_step": gather_bytes,
                    "what": "scatter_items + pinned-host conditioning H2D + sample + decode + uint8 frames (mudg_postdecode) + NCCL gather to rank 0 + D2H"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "scope": "whole path (UNet forwards + VAE decode + sampler glue), wall time of the timed region",
                         "achieved": path_tf, "peak": sus, "unit": "TFLOP/s", "frac": path_tf / sus, "frac_of_burst": path_tf / burst,
                         "peak_source": how, "algorithmic_tflop_per_clip": flops_clip / 1e12, "traffic": None,
                         "instrumented_step_ms": ms_prof,
                         "how": "kernels[]: CUDA-event pair around every launch of one extra (eager) clip after the timed region; "
                                "algorithmic FLOPs / HBM bytes per launch as DESIGN.md section 4 states them",
                         "kernels": kernels, "ncu": ncu},
        }
        if parity:
            res["parity"] = parity
        if extra:
            res["extra"] = extra
        if not args.no_cpu_baseline and world == 1:
            try:
                res["cpu_baseline"] = cpu_baseline_leg(model, args.config)
            except Exception as e:
                res["cpu_baseline"] = {"error": repr(e)[:200]}
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
