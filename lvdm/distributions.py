"""Diagonal Gaussian posterior of the KL-VAE (lvdm/distributions.py:24-61).  sample() draws on the CPU
generator and moves the noise to the parameters' device, exactly like the reference (:35-40)."""
import torch


class DiagonalGaussianDistribution:
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape)
        return self.mean + self.std * noise.to(device=self.parameters.device)

    def mode(self):
        return self.mean
