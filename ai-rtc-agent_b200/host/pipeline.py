"""Drop-in for the reference's `lib/pipeline.py:StreamDiffusionPipeline` (lib/pipeline.py:17-96): same
module constants, constructor, attributes and methods, so agent.py:23,423 and lib/tracks.py:24,38 run on
it unchanged.

Frames: the reference accepts `nvcv.Tensor` (NVDEC path) or `av.VideoFrame` (software decode).  Neither
package exists offline, so frames are recognised structurally: anything exposing
`__cuda_array_interface__` / `.cuda()` (nvcv.Tensor, torch.Tensor) is a GPU u8 NHWC frame, anything with
`.to_ndarray` is an av.VideoFrame; everything else raises Exception("invalid frame type") like
lib/pipeline.py:51-52.

Fast path: a GPU u8 frame goes through ONE engine call (pre + encode + UNet + decode + post fused,
u8 NHWC in -> u8 NCHW out, nothing leaves HBM).  preprocess / predict / postprocess remain callable
separately with the reference's tensor contracts."""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch

from .wrapper import StreamDiffusionWrapper

DEFAULT_PROMPT = "fireworks in the night sky"
DEFAULT_T_INDEX_LIST = [18, 26, 35, 45]
DEFAULT_NUM_INFERENCE_STEPS = 50
DEFAULT_GUIDANCE_SCALE = 0.0


def _is_video_frame(frame) -> bool:
    return hasattr(frame, "to_ndarray") and hasattr(frame, "pts")


def _is_gpu_frame(frame) -> bool:
    if isinstance(frame, torch.Tensor):
        return frame.is_cuda
    return hasattr(frame, "__cuda_array_interface__") or (hasattr(frame, "cuda") and hasattr(frame, "layout"))


def _as_torch_u8_nhwc(frame, device) -> torch.Tensor:
    if isinstance(frame, torch.Tensor):
        t = frame
    elif hasattr(frame, "cuda") and not hasattr(frame, "__cuda_array_interface__"):
        t = torch.as_tensor(frame.cuda(), device=device)  # nvcv.Tensor
    else:
        t = torch.as_tensor(frame, device=device)
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dtype != torch.uint8 or t.shape[-1] != 3:
        raise Exception("invalid frame type")
    return t


class StreamDiffusionPipeline:
    def __init__(self, model_id: str, t_index_list: Optional[List[int]] = None, width: int = 512, height: int = 512,
                 prompt: str = DEFAULT_PROMPT):
        self.prompt = prompt
        self.t_index_list = list(t_index_list) if t_index_list is not None else DEFAULT_T_INDEX_LIST
        self.device = "cuda"
        self.model = StreamDiffusionWrapper(
            model_id_or_path=model_id,
            device=self.device,
            dtype=torch.float16,
            t_index_list=self.t_index_list,
            frame_buffer_size=1,
            width=width,
            height=height,
            use_lcm_lora=True,
            output_type="pt",
            mode="img2img",
            use_denoising_batch=True,
            use_tiny_vae=True,
            cfg_type="self",
            engine_dir=os.getenv("TRT_ENGINES_CACHE", "./models/engines"),
        )
        self.model.prepare(prompt=self.prompt, num_inference_steps=DEFAULT_NUM_INFERENCE_STEPS,
                           guidance_scale=DEFAULT_GUIDANCE_SCALE)

    def update_prompt(self, prompt: str):
        self.model.stream.update_prompt(prompt)

    def update_t_index_list(self, t_index_list: List[int]):
        self.model.update_t_index_list(t_index_list)

    # ---- reference-shaped stages ----------------------------------------------------------------------
    def preprocess(self, frame) -> torch.Tensor:
        """-> (3,H,W) float32 in [0,1] on the GPU (lib/pipeline.py:50-67)."""
        if not _is_gpu_frame(frame) and not _is_video_frame(frame):
            raise Exception("invalid frame type")
        if _is_video_frame(frame):
            t = torch.from_numpy(frame.to_ndarray(format="rgb24")).unsqueeze(0).to(self.device)
        else:
            t = _as_torch_u8_nhwc(frame, self.device)
        return (t.to(torch.float32) * (1.0 / 255.0)).permute(0, 3, 1, 2).squeeze(0)

    def predict(self, frame: torch.Tensor) -> torch.Tensor:
        return self.model(image=frame)

    def postprocess(self, frame: torch.Tensor) -> torch.Tensor:
        """(3,H,W) in [0,1] -> (1,3,H,W) uint8; the cast truncates (lib/pipeline.py:72-74)."""
        return frame.mul(255.0).clamp_(0, 255).to(torch.uint8)[None]

    def __call__(self, frame):
        if not _is_gpu_frame(frame) and not _is_video_frame(frame):
            raise Exception("invalid frame type")
        if _is_video_frame(frame):
            rgb = torch.from_numpy(frame.to_ndarray(format="rgb24")).unsqueeze(0).to(self.device)
        else:
            rgb = _as_torch_u8_nhwc(frame, self.device)
        post_output = self.model.stream.step_u8(rgb)

        if not os.getenv("NVENC"):
            # software-encode branch (lib/pipeline.py:83-94): hand back an av.VideoFrame with the input's timing
            try:
                import av
            except ImportError as exc:
                raise RuntimeError("NVENC is unset, so an av.VideoFrame must be returned, but PyAV is not installed; "
                                   "set NVENC=1 to receive the CUDA tensor") from exc
            assert _is_video_frame(frame)
            hwc = post_output.cpu().permute(0, 2, 3, 1).squeeze(0).numpy()
            out = av.VideoFrame.from_ndarray(np.ascontiguousarray(hwc))
            out.pts = frame.pts
            out.time_base = frame.time_base
            return out
        return post_output
