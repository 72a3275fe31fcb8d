"""Schedules and small tensor helpers of the sampler, restated from the reference's formulas
(lvdm/models/utils_diffusion.py).  Host-side NumPy/f64 like the reference."""
import math

import numpy as np
import torch


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """[cos | sin] sinusoid (utils_diffusion.py:8-28)."""
    if repeat_only:
        return timesteps[:, None].expand(-1, dim)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """utils_diffusion.py:31-53.  float64 torch arithmetic in the reference's order (np.linspace / np.cos differ from
    torch's in the last ulp, and the tables are compared bit for bit); returns a numpy array like the reference."""
    f64 = torch.float64
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=f64) ** 2
    elif schedule == "cosine":
        t = torch.arange(n_timestep + 1, dtype=f64) / n_timestep + cosine_s
        a = torch.cos(t / (1 + cosine_s) * np.pi / 2).pow(2)
        a = a / a[0]
        betas = torch.from_numpy(np.clip((1 - a[1:] / a[:-1]).numpy(), 0, 0.999))
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=f64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=f64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


def rescale_zero_terminal_snr(betas):
    """Zero terminal SNR (arXiv 2305.08891 alg. 1; utils_diffusion.py:112-144)."""
    s = np.sqrt(np.cumprod(1.0 - betas))
    s0, sT = s[0], s[-1]
    s = (s - sT) * (s0 / (s0 - sT))
    abar = s ** 2
    return 1.0 - np.concatenate([abar[:1], abar[1:] / abar[:-1]])


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """utils_diffusion.py:56-76."""
    if ddim_discr_method == "uniform":
        steps = np.arange(0, num_ddpm_timesteps, num_ddpm_timesteps // num_ddim_timesteps) + 1
    elif ddim_discr_method == "uniform_trailing":
        c = num_ddpm_timesteps / num_ddim_timesteps
        steps = np.flip(np.round(np.arange(num_ddpm_timesteps, 0, -c))).astype(np.int64) - 1
    elif ddim_discr_method == "quad":
        steps = (np.linspace(0, np.sqrt(num_ddpm_timesteps * 0.8), num_ddim_timesteps) ** 2).astype(int) + 1
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps}")
    return steps


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """(sigmas, alphas, alphas_prev) as f64/f32 NumPy (utils_diffusion.py:79-91).  `alphacums` is the model's fp32
    table.  The reference's mixed ndarray/Tensor expression evaluates 1/(1-alpha_t) as an fp32 reciprocal and the
    rest in f64; reproduced so sigma_t is bit-identical."""
    ac = alphacums.detach().cpu() if torch.is_tensor(alphacums) else torch.as_tensor(alphacums)
    alphas = ac[ddim_timesteps].to(torch.float32).numpy()
    alphas_prev = np.asarray([float(ac[0])] + ac[ddim_timesteps[:-1]].tolist())
    recip = (np.float32(1.0) / (np.float32(1.0) - alphas)).astype(np.float64)
    sigmas = eta * np.sqrt(recip * (1 - alphas_prev) * (1 - alphas.astype(np.float64) / alphas_prev))
    if verbose:
        print(f"ddim alphas {alphas}; alphas_prev {alphas_prev}; eta {eta} -> sigmas {sigmas}")
    return sigmas, alphas, alphas_prev


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """utils_diffusion.py:147-158."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=dims, keepdim=True)
    return guidance_rescale * (noise_cfg * (std_text / std_cfg)) + (1 - guidance_rescale) * noise_cfg
