"""Where does a benchmark clip spend its time?  sampler.sample (50 steps) and decode_first_stage timed separately, per-step
wall/GPU times, through the same public API as bench.py."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                   # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    model = bench.build_model("mdm1024", dev)
    from lvdm.models.samplers.ddim import DDIMSampler
    sampler = DDIMSampler(model)
    T, h, w = 16, 72, 128
    g = torch.Generator().manual_seed(123)
    ctx = torch.randn(1, 77 + 16 * T, 1024, generator=g).to(dev)
    uc = torch.randn(1, 77 + 16 * T, 1024, generator=g).to(dev)
    cat = (0.5 * torch.randn(1, 8, T, h, w, generator=g)).to(dev)
    label = torch.zeros(1, 1, dtype=torch.long, device=dev)
    fs = torch.full((1,), 10, dtype=torch.long, device=dev)
    times = []
    def cb(i):
        e = torch.cuda.Event(enable_timing=True); e.record(); times.append((time.time(), e))
    for rep in range(3):
        times.clear()
        torch.manual_seed(123)
        cond = {"c_crossattn": [ctx], "c_concat": [cat]}
        ucd = {"c_crossattn": [uc], "c_concat": [cat]}
        torch.cuda.synchronize()
        t0 = time.time()
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            z, _ = sampler.sample(S=50, conditioning=cond, batch_size=1, shape=[4, T, h, w], verbose=False,
                                  unconditional_guidance_scale=7.5, unconditional_conditioning=ucd, eta=1.0, cfg_img=None,
                                  mask=None, x0=None, fs=fs, timestep_spacing="uniform_trailing", guidance_rescale=0.7,
                                  sparse_x=None, class_label=label, unconditional_conditioning_img_nonetext=None, callback=cb)
            e1 = torch.cuda.Event(enable_timing=True); e1.record()
            t_issue = time.time() - t0
            fr = model.decode_first_stage(z)
        e2 = torch.cuda.Event(enable_timing=True); e2.record()
        torch.cuda.synchronize()
        steps = [times[i][1].elapsed_time(times[i + 1][1]) for i in range(len(times) - 1)]
        print(f"rep {rep}: sample {e0.elapsed_time(e1):.0f} ms (host issue {t_issue * 1e3:.0f} ms), decode {e1.elapsed_time(e2):.0f} ms; "
              f"per-step GPU ms: first {e0.elapsed_time(times[0][1]):.1f}, min {min(steps):.1f} median {sorted(steps)[len(steps) // 2]:.1f} "
              f"max {max(steps):.1f}", flush=True)


if __name__ == "__main__":
    main()
