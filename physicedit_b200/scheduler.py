"""Flow-matching scheduler -- host-side bookkeeping, bit-exact with the reference.

Mirrors DiffSynth-Studio/diffsynth/schedulers/flow_match.py (FlowMatchScheduler) for the argument set the
PhysicEdit pipeline uses (qwen_image_physical.py:192): the sigma / timestep tables are fp32 CPU tensors
built with the same torch ops in the same order, so indices and values are identical.  The Euler update
itself runs on the device in `pe_cfg_euler_step`; `dsigma(progress_id)` hands it the fp32 scalar.
"""
import math

import torch


class FlowMatchScheduler:
    """Only the configuration PhysicEdit constructs (qwen_image_physical.py:192: extra_one_step, exponential shift with a
    terminal stretch) is implemented; the other modes of the reference class (plain `shift`, inverse / reversed sigmas) are
    rejected instead of carried along untested."""

    def __init__(self, num_inference_steps=100, num_train_timesteps=1000, sigma_max=1.0, sigma_min=0.0, extra_one_step=True,
                 exponential_shift=True, exponential_shift_mu=0.8, shift_terminal=0.02, **unsupported):
        bad = {k: v for k, v in unsupported.items() if v not in (None, False) and k != "shift"}
        if bad or not (extra_one_step and exponential_shift) or shift_terminal is None:
            raise NotImplementedError(f"FlowMatchScheduler: only the Qwen-Image configuration is implemented (got {bad or 'a non-exponential schedule'})")
        self.num_train_timesteps = num_train_timesteps
        self.sigma_max, self.sigma_min = sigma_max, sigma_min
        self.exponential_shift_mu = exponential_shift_mu
        self.shift_terminal = shift_terminal
        self.inverse_timesteps = self.reverse_sigmas = False
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False, dynamic_shift_len=None,
                      exponential_shift_mu=None, **_):
        """flow_match.py:34-69 for this configuration.  The torch ops and their order are the bit-exactness contract (fp32 CPU
        tensors; `math.exp(mu)` in double, folded into the tensor expression exactly as there)."""
        top = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        grid = torch.linspace(top, self.sigma_min, num_inference_steps + 1)[:-1]           # extra_one_step: N+1 points, last dropped
        if exponential_shift_mu is not None:
            mu = exponential_shift_mu
        elif dynamic_shift_len is not None:
            mu = self.calculate_shift(dynamic_shift_len)
        else:
            mu = self.exponential_shift_mu
        shifted = math.exp(mu) / (math.exp(mu) + (1 / grid - 1))
        gap = 1 - shifted                                                                    # stretch so that the last sigma is shift_terminal
        self.sigmas = 1 - (gap / (gap[-1] / (1 - self.shift_terminal)))
        self.timesteps = self.sigmas * self.num_train_timesteps
        self.training = bool(training)
        if training:                                                                         # :62-66 bell-shaped loss weights, sum = N
            bell = torch.exp(-2 * ((self.timesteps - num_inference_steps / 2) / num_inference_steps) ** 2)
            bell = bell - bell.min()
            self.linear_timesteps_weights = bell * (num_inference_steps / bell.sum())

    def _timestep_id(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.cpu()
        return torch.argmin((self.timesteps - timestep).abs())

    def dsigma(self, timestep, to_final=False) -> torch.Tensor:
        """(sigma_next - sigma) of FlowMatchScheduler.step (flow_match.py:72-81) as a 0-dim fp32 CPU tensor."""
        tid = self._timestep_id(timestep)
        sigma = self.sigmas[tid]
        if to_final or tid + 1 >= len(self.timesteps):
            sigma_ = 0
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
        """:114-125: mu is linear in the number of latent tokens (0.5 at 256, 0.9 at 8192, extrapolating beyond).  Same double
        arithmetic, same order: slope, intercept, then len * slope + intercept."""
        slope = (max_shift - base_shift) / (max_seq_len - base_seq_len)
        intercept = base_shift - slope * base_seq_len
        return image_seq_len * slope + intercept
