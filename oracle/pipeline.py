"""Restatement of the reference's own (in-repo) pre/post-processing around the model call:
lib/pipeline.py:50-67 (preprocess), :72-74 (postprocess), lib/wrapper.py:368-387 +
streamdiffusion.image_utils.postprocess_image (denormalise).  Test infrastructure only."""
from __future__ import annotations

import torch

from .stream import StreamOracle


def preprocess(frame_u8_nhwc: torch.Tensor) -> torch.Tensor:
    """lib/pipeline.py:61-65: u8 NHWC (1,H,W,3) -> f32 /255 -> NCHW -> squeeze(0) => (3,H,W) in [0,1]."""
    x = frame_u8_nhwc.to(torch.float32) * (1.0 / 255.0)
    return x.permute(0, 3, 1, 2).squeeze(0)


def denormalize_pt(image: torch.Tensor) -> torch.Tensor:
    """image_utils.postprocess_image(.., "pt") for frame_buffer_size == 1 (lib/wrapper.py:384-387):
    (x/2 + 0.5).clamp(0,1), first item."""
    return (image / 2 + 0.5).clamp(0, 1)[0]


def postprocess(frame: torch.Tensor) -> torch.Tensor:
    """lib/pipeline.py:72-74: (x*255).clamp(0,255).to(uint8).unsqueeze(0); the float->u8 cast
    truncates toward zero."""
    return (frame * 255.0).clamp(0, 255).to(dtype=torch.uint8).unsqueeze(0)


def frame_to_u8(stream: StreamOracle, frame_u8_nhwc: torch.Tensor, fp16_tail: bool = True) -> torch.Tensor:
    """One StreamDiffusionPipeline.__call__ (lib/pipeline.py:76-96, NVENC branch): u8 NHWC in,
    (1,3,H,W) u8 NCHW out.  fp16_tail reproduces the dtype of the reference's tail: the model output
    is fp16, so denormalise / *255 / truncation happen on the fp16 grid."""
    out = stream(preprocess(frame_u8_nhwc))
    if fp16_tail:
        out = out.to(torch.float16)
    return postprocess(denormalize_pt(out))
