"""fp32 restatement of `streamdiffusion.StreamDiffusion` (yondonfu/StreamDiffusion@deepstream,
requirements.txt:14 -- not vendored) as driven by lib/wrapper.py:197-234 (prepare), :302-343 (img2img),
:389-407 (update_t_index_list), plus diffusers' LCMScheduler tables and VaeImageProcessor.preprocess.
SURVEY.md Appendix A.1 / A.4.  Text encoding is outside the per-frame path: prompts enter as
embeddings (B,77,D).  Test infrastructure only (see oracle/__init__)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import taesd, unet

NUM_TRAIN_TIMESTEPS = 1000
BETA_START, BETA_END = 0.00085, 0.012   # scaled_linear schedule shared by SD-1.5 / SD-2.1 / SD-Turbo
LCM_ORIGINAL_STEPS = 50
SIGMA_DATA = 0.5
TIMESTEP_SCALING = 10.0


def alphas_cumprod() -> torch.Tensor:
    betas = torch.linspace(BETA_START ** 0.5, BETA_END ** 0.5, NUM_TRAIN_TIMESTEPS, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def lcm_timesteps(num_inference_steps: int) -> List[int]:
    """LCMScheduler.set_timesteps: origin = arange(1, 51)*20 - 1, reversed, every k-th, first N."""
    k = NUM_TRAIN_TIMESTEPS // LCM_ORIGINAL_STEPS
    origin = [i * k - 1 for i in range(1, LCM_ORIGINAL_STEPS + 1)]
    skip = len(origin) // num_inference_steps
    return origin[::-1][::skip][:num_inference_steps]


def boundary_scalings(timestep: int):
    """LCMScheduler.get_scalings_for_boundary_condition_discrete."""
    s = timestep * TIMESTEP_SCALING
    c_skip = SIGMA_DATA ** 2 / (s ** 2 + SIGMA_DATA ** 2)
    c_out = s / (s ** 2 + SIGMA_DATA ** 2) ** 0.5
    return c_skip, c_out


def image_preprocess(x: torch.Tensor, height: int, width: int, assume_unit_range: bool = False) -> torch.Tensor:
    """VaeImageProcessor.preprocess for tensor input: 4-D, nearest resize if needed, 2x-1 unless the
    image already has negative values."""
    if x.dim() == 3:
        x = x.unsqueeze(0)
    if x.shape[-2] != height or x.shape[-1] != width:
        x = F.interpolate(x, size=(height, width))
    if not assume_unit_range and x.min() < 0:
        return x
    return 2.0 * x - 1.0


class StreamOracle:
    def __init__(self, unet_sd: Dict[str, torch.Tensor], unet_cfg: unet.UNetConfig,
                 vae_sd: Dict[str, torch.Tensor], t_index_list: List[int], width: int = 512,
                 height: int = 512, do_add_noise: bool = True, frame_buffer_size: int = 1,
                 cfg_type: str = "self", run_dead_code: bool = False):
        self.unet_sd, self.cfg, self.vae_sd = unet_sd, unet_cfg, vae_sd
        self.t_list = list(t_index_list)
        self.width, self.height = width, height
        self.latent_h, self.latent_w = height // 8, width // 8
        self.do_add_noise = do_add_noise
        self.frame_bff_size = frame_buffer_size
        self.denoising_steps_num = len(t_index_list)
        self.batch_size = self.denoising_steps_num * frame_buffer_size
        self.cfg_type = cfg_type
        self.use_denoising_batch = True
        self.run_dead_code = run_dead_code
        self.guidance_scale = 1.0
        self.x_t_latent_buffer: Optional[torch.Tensor] = None
        self.last = {}   # intermediates of the last call, for parity debugging

    # ---- StreamDiffusion.prepare
    def prepare(self, prompt_embeds: torch.Tensor, num_inference_steps: int = 50, guidance_scale: float = 1.2,
                delta: float = 1.0, seed: int = 2, init_noise: Optional[torch.Tensor] = None) -> None:
        gen = torch.Generator().manual_seed(seed)
        T, Fb = self.denoising_steps_num, self.frame_bff_size
        shape = (4, self.latent_h, self.latent_w)
        self.x_t_latent_buffer = torch.zeros(((T - 1) * Fb, *shape)) if T > 1 else None
        self.guidance_scale = 1.0 if self.cfg_type == "none" else guidance_scale
        self.delta = delta
        self.prompt_embeds = prompt_embeds.float().reshape(1, prompt_embeds.shape[-2], -1).repeat(self.batch_size, 1, 1)
        self.timesteps = lcm_timesteps(num_inference_steps)
        self.sub_timesteps = [self.timesteps[t] for t in self.t_list]
        self.sub_timesteps_tensor = torch.tensor(self.sub_timesteps, dtype=torch.long).repeat_interleave(Fb)
        self.init_noise = (init_noise.float().clone() if init_noise is not None
                           else torch.randn((self.batch_size, *shape), generator=gen))
        self.stock_noise = torch.zeros_like(self.init_noise)
        cs = [boundary_scalings(t) for t in self.sub_timesteps]
        self.c_skip = torch.tensor([c[0] for c in cs]).view(T, 1, 1, 1).repeat_interleave(Fb, 0)
        self.c_out = torch.tensor([c[1] for c in cs]).view(T, 1, 1, 1).repeat_interleave(Fb, 0)
        ac = alphas_cumprod()
        self.alpha_prod_t_sqrt = torch.stack([ac[t].sqrt() for t in self.sub_timesteps]).view(T, 1, 1, 1).repeat_interleave(Fb, 0)
        self.beta_prod_t_sqrt = torch.stack([(1 - ac[t]).sqrt() for t in self.sub_timesteps]).view(T, 1, 1, 1).repeat_interleave(Fb, 0)

    def update_prompt_embeds(self, prompt_embeds: torch.Tensor) -> None:
        pe = prompt_embeds.reshape(1, prompt_embeds.shape[-2], -1).repeat(self.batch_size, 1, 1)
        self.prompt_embeds = pe.to(device=self.device, dtype=self.dtype)

    # ---- the same restatement on another device / dtype (oracle/torch_gpu.py: fp32 or fp16 on the GPU through torch's
    # library kernels).  Everything the per-frame path touches moves; the math is unchanged.
    device = torch.device("cpu")
    dtype = torch.float32
    static_buffers = False      # True: stream state is updated in place (CUDA-graph capture of the frame)
    assume_unit_range = False   # True: skip the `image.min() < 0` host sync of VaeImageProcessor (CUDA-graph capture)

    def to(self, device, dtype: torch.dtype = torch.float32) -> "StreamOracle":
        self.device, self.dtype = torch.device(device), dtype
        mv = lambda t: t.to(device=self.device, dtype=dtype)
        self.unet_sd = {k: mv(v) for k, v in self.unet_sd.items()}
        self.vae_sd = {k: mv(v) for k, v in self.vae_sd.items()}
        for name in ("prompt_embeds", "init_noise", "stock_noise", "c_skip", "c_out", "alpha_prod_t_sqrt", "beta_prod_t_sqrt",
                     "x_t_latent_buffer"):
            v = getattr(self, name, None)
            if v is not None:
                setattr(self, name, mv(v))
        if hasattr(self, "sub_timesteps_tensor"):
            self.sub_timesteps_tensor = self.sub_timesteps_tensor.to(self.device)
        return self

    # ---- lib/wrapper.py:389-407: rebuilds sub_timesteps only; alpha/beta/c_skip/c_out keep the
    # values prepare() derived from the previous list (reference quirk, reproduced on purpose)
    def update_t_index_list(self, t_index_list: List[int]) -> None:
        if t_index_list == self.t_list:
            return
        self.t_list = list(t_index_list)
        self.sub_timesteps = [self.timesteps[t] for t in t_index_list]
        self.sub_timesteps_tensor = torch.tensor(self.sub_timesteps, dtype=torch.long).repeat_interleave(self.frame_bff_size).to(self.device)

    # ---- per-frame
    def scheduler_step_batch(self, eps: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        f_theta = (x - self.beta_prod_t_sqrt * eps) / self.alpha_prod_t_sqrt
        return self.c_out * f_theta + self.c_skip * x

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        z = taesd.encode(self.vae_sd, image) * taesd.SCALING_FACTOR
        return self.alpha_prod_t_sqrt[0] * z + self.beta_prod_t_sqrt[0] * self.init_noise[0]

    def unet_step(self, x: torch.Tensor):
        eps = unet.unet_forward(self.unet_sd, self.cfg, x, self.sub_timesteps_tensor, self.prompt_embeds)
        x0 = self.scheduler_step_batch(eps, x)
        if self.run_dead_code and self.cfg_type in ("self", "initialize"):
            # result is only consumed when guidance_scale > 1 (never on the reference path,
            # lib/pipeline.py:14 passes 0.0); kept for completeness
            scaled = self.beta_prod_t_sqrt * self.stock_noise
            delta_x = self.scheduler_step_batch(eps, scaled)
            one = torch.ones_like(self.alpha_prod_t_sqrt[0:1])
            a_next = torch.cat([self.alpha_prod_t_sqrt[1:], one], 0)
            b_next = torch.cat([self.beta_prod_t_sqrt[1:], one], 0)
            init_roll = torch.cat([self.init_noise[1:], self.init_noise[0:1]], 0)
            self.stock_noise = init_roll + a_next * delta_x / b_next
        return x0, eps

    def predict_x0_batch(self, x_t: torch.Tensor) -> torch.Tensor:
        T = self.denoising_steps_num
        if T > 1:
            x = torch.cat([x_t, self.x_t_latent_buffer], 0)
            self.stock_noise = torch.cat([self.init_noise[0:1], self.stock_noise[:-1]], 0)
        else:
            x = x_t
        x0_batch, eps = self.unet_step(x)
        self.last.update(unet_in=x, eps=eps, x0_batch=x0_batch)
        if T > 1:
            out = x0_batch[-1:]
            if self.do_add_noise:
                nxt = self.alpha_prod_t_sqrt[1:] * x0_batch[:-1] + self.beta_prod_t_sqrt[1:] * self.init_noise[1:]
            else:
                nxt = self.alpha_prod_t_sqrt[1:] * x0_batch[:-1]
            if self.static_buffers:
                self.x_t_latent_buffer.copy_(nxt)   # CUDA-graph replay needs the state at a fixed address
            else:
                self.x_t_latent_buffer = nxt
            return out
        return x0_batch

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """x: (3,H',W') or (1,3,H',W') float in [0,1] -> (1,3,H,W) float, roughly [-1,1]."""
        img = image_preprocess(x.to(device=self.device, dtype=self.dtype), self.height, self.width, self.assume_unit_range)
        x_t = self.encode_image(img)
        x0 = self.predict_x0_batch(x_t)
        out = taesd.decode(self.vae_sd, x0 / taesd.SCALING_FACTOR)
        self.last.update(x_t=x_t, x0=x0, image=out)
        return out
