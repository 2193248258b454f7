"""Regenerates tests/golden/*.npz from the CPU oracle (python tests/golden/make_golden.py).

The reference has no golden vectors and its compute dependencies cannot be imported here,
so these fixtures pin the ORACLE's behaviour (guarding it against regressions) and give the GPU tests a
committed target that does not require running the oracle: seeded tiny-width models, 64x64 frames."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pipeline as opipe  # noqa: E402
from oracle import stream as ostream  # noqa: E402
from oracle import unet as ounet  # noqa: E402
from oracle import weights as ow  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"tiny_turbo_T1": (True, [32]), "tiny_turbo_T2": (True, [20, 40]), "tiny_sd15_T4": (False, [18, 26, 35, 45])}
HW, NFRAMES = 128, 6


def run_case(turbo, t_index_list):
    torch.manual_seed(0)
    torch.set_num_threads(4)
    cfg = ounet.tiny_config(turbo)
    usd, vsd = ow.to_float(ow.make_unet_weights(cfg)), ow.to_float(ow.make_taesd_weights())
    orc = ostream.StreamOracle(usd, cfg, vsd, t_index_list, HW, HW)
    emb = ow.make_prompt_embeds(cfg.cross_attention_dim)
    orc.prepare(emb.float(), guidance_scale=0.0, seed=2)
    # the engine receives fp16 noise: pin the oracle to the same rounded values
    orc.init_noise = orc.init_noise.half().float()
    outs, eps = [], []
    for i in range(NFRAMES):
        outs.append(opipe.frame_to_u8(orc, ow.make_frame(HW, HW, seed=i)).numpy())
        eps.append(orc.last["eps"].numpy().astype(np.float32))
    return {"u8": np.concatenate(outs, 0), "eps": np.stack(eps, 0), "sub_timesteps": np.array(orc.sub_timesteps),
            "alpha": orc.alpha_prod_t_sqrt.flatten().numpy(), "beta": orc.beta_prod_t_sqrt.flatten().numpy(),
            "c_skip": orc.c_skip.flatten().numpy(), "c_out": orc.c_out.flatten().numpy()}


if __name__ == "__main__":
    for name, (turbo, tl) in CASES.items():
        d = run_case(turbo, tl)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: v.shape for k, v in d.items()})
