"""Prompt -> (1,77,D) embedding.  The reference keeps the CLIP text encoder in torch (lib/wrapper.py:468-473)
and calls it only on prepare / update_prompt -- it is off the per-frame path, so it stays a torch module
here too (plumbing).  Without a checkpoint on disk (download.py needs network) a deterministic synthetic
encoder keeps the pipeline runnable for benchmarks and tests."""
from __future__ import annotations

import hashlib
import logging
import os
from typing import Optional

import torch

logger = logging.getLogger(__name__)


class SyntheticPromptEncoder:
    """Deterministic stand-in: seeds a CPU generator from sha256(prompt)."""

    def __init__(self, dim: int, tokens: int = 77):
        self.dim, self.tokens = dim, tokens

    def __call__(self, prompt: str) -> torch.Tensor:
        seed = int.from_bytes(hashlib.sha256(prompt.encode("utf-8")).digest()[:8], "little") % (2 ** 63)
        g = torch.Generator().manual_seed(seed)
        return torch.randn((1, self.tokens, self.dim), generator=g).to(torch.float16)


class ClipPromptEncoder:
    """CLIPTextModel + CLIPTokenizer from a local diffusers model directory (text_encoder/, tokenizer/)."""

    def __init__(self, model_dir: str, device: str = "cuda"):
        from transformers import CLIPTextModel, CLIPTokenizer
        self.tokenizer = CLIPTokenizer.from_pretrained(os.path.join(model_dir, "tokenizer"))
        self.model = CLIPTextModel.from_pretrained(os.path.join(model_dir, "text_encoder"),
                                                   torch_dtype=torch.float16).to(device).eval()
        self.device = device

    @torch.no_grad()
    def __call__(self, prompt: str) -> torch.Tensor:
        ids = self.tokenizer(prompt, padding="max_length", max_length=self.tokenizer.model_max_length,
                             truncation=True, return_tensors="pt").input_ids.to(self.device)
        return self.model(ids)[0].to(torch.float16)


def make_prompt_encoder(model_dir: Optional[str], dim: int, device: str = "cuda", allow_synthetic: bool = False):
    """A checkpoint directory with a text_encoder/ must load (a real model fed hash-seeded embeddings renders garbage for
    every prompt: failing is the only safe behaviour).  The synthetic encoder is for synthetic weights only."""
    if model_dir and os.path.isdir(os.path.join(model_dir, "text_encoder")):
        enc = ClipPromptEncoder(model_dir, device)
        got = enc.model.config.hidden_size
        if got != dim:
            raise ValueError(f"text encoder under {model_dir} has hidden size {got}, the UNet cross-attention expects {dim}")
        return enc
    if model_dir and not allow_synthetic:
        raise FileNotFoundError(f"{model_dir} has no text_encoder/: cannot encode prompts for a real checkpoint "
                                f"(set B200SD_SYNTHETIC_WEIGHTS=1 to run with synthetic embeddings)")
    return SyntheticPromptEncoder(dim)
