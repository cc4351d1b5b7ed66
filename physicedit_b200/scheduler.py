"""Flow-matching scheduler -- host-side bookkeeping, bit-exact with the reference.

Mirrors DiffSynth-Studio/diffsynth/schedulers/flow_match.py (FlowMatchScheduler) for the argument set the
PhysicEdit pipeline uses (qwen_image_physical.py:192): the sigma / timestep tables are fp32 CPU tensors
built with the same torch ops in the same order, so indices and values are identical.  The Euler update
itself runs on the device in `pe_cfg_euler_step`; `dsigma(progress_id)` hands it the fp32 scalar.
"""
import math

import torch


class FlowMatchScheduler:
    def __init__(self, num_inference_steps=100, num_train_timesteps=1000, shift=3.0, sigma_max=1.0, sigma_min=0.003 / 1.002,
                 inverse_timesteps=False, extra_one_step=False, reverse_sigmas=False, exponential_shift=False,
                 exponential_shift_mu=None, shift_terminal=None):
        self.num_train_timesteps = num_train_timesteps
        self.shift = shift
        self.sigma_max, self.sigma_min = sigma_max, sigma_min
        self.inverse_timesteps, self.extra_one_step, self.reverse_sigmas = inverse_timesteps, extra_one_step, reverse_sigmas
        self.exponential_shift, self.exponential_shift_mu = exponential_shift, exponential_shift_mu
        self.shift_terminal = shift_terminal
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False, shift=None, dynamic_shift_len=None,
                      exponential_shift_mu=None):
        if shift is not None:
            self.shift = shift
        sigma_start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        if self.extra_one_step:
            self.sigmas = torch.linspace(sigma_start, self.sigma_min, num_inference_steps + 1)[:-1]
        else:
            self.sigmas = torch.linspace(sigma_start, self.sigma_min, num_inference_steps)
        if self.inverse_timesteps:
            self.sigmas = torch.flip(self.sigmas, dims=[0])
        if self.exponential_shift:
            if exponential_shift_mu is not None:
                mu = exponential_shift_mu
            elif dynamic_shift_len is not None:
                mu = self.calculate_shift(dynamic_shift_len)
            else:
                mu = self.exponential_shift_mu
            self.sigmas = math.exp(mu) / (math.exp(mu) + (1 / self.sigmas - 1))
        else:
            self.sigmas = self.shift * self.sigmas / (1 + (self.shift - 1) * self.sigmas)
        if self.shift_terminal is not None:
            one_minus_z = 1 - self.sigmas
            scale_factor = one_minus_z[-1] / (1 - self.shift_terminal)
            self.sigmas = 1 - (one_minus_z / scale_factor)
        if self.reverse_sigmas:
            self.sigmas = 1 - self.sigmas
        self.timesteps = self.sigmas * self.num_train_timesteps
        if training:
            x = self.timesteps
            y = torch.exp(-2 * ((x - num_inference_steps / 2) / num_inference_steps) ** 2)
            y_shifted = y - y.min()
            self.linear_timesteps_weights = y_shifted * (num_inference_steps / y_shifted.sum())
            self.training = True
        else:
            self.training = False

    def _timestep_id(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.cpu()
        return torch.argmin((self.timesteps - timestep).abs())

    def dsigma(self, timestep, to_final=False) -> torch.Tensor:
        """(sigma_next - sigma) of FlowMatchScheduler.step (flow_match.py:72-81) as a 0-dim fp32 CPU tensor."""
        tid = self._timestep_id(timestep)
        sigma = self.sigmas[tid]
        if to_final or tid + 1 >= len(self.timesteps):
            sigma_ = 1 if (self.inverse_timesteps or self.reverse_sigmas) else 0
        else:
            sigma_ = self.sigmas[tid + 1]
        return sigma_ - sigma

    def step(self, model_output, timestep, sample, to_final=False, **kwargs):
        """Same contract as the reference; on CUDA bf16 tensors the update runs in pe_cfg_euler_step."""
        ds = self.dsigma(timestep, to_final)
        if sample.is_cuda and sample.dtype == torch.bfloat16:
            from .native import Native
            out = sample.clone()
            Native.get(sample.device.index or 0).cfg_euler_step(out, model_output.contiguous(), None, 1.0, float(ds))
            return out
        return sample + model_output * ds

    def return_to_timestep(self, timestep, sample, sample_stablized):
        sigma = self.sigmas[self._timestep_id(timestep)]
        return (sample - sample_stablized) / sigma

    def add_noise(self, original_samples, noise, timestep):
        sigma = self.sigmas[self._timestep_id(timestep)]
        return (1 - sigma) * original_samples + sigma * noise

    def training_target(self, sample, noise, timestep):
        return noise - sample

    def training_weight(self, timestep):
        tid = torch.argmin((self.timesteps - timestep.to(self.timesteps.device)).abs())
        return self.linear_timesteps_weights[tid]

    def calculate_shift(self, image_seq_len, base_seq_len: int = 256, max_seq_len: int = 8192, base_shift: float = 0.5,
                        max_shift: float = 0.9):
        m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
        b = base_shift - m * base_seq_len
        return image_seq_len * m + b
