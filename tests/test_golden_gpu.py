"""GPU path vs the committed golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the
CPU oracle): needs neither /root/reference nor an oracle run on the GPU box.  Tolerances as in test_engine_gpu.py."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,turbo,tl", [("tiny_turbo_T1", True, [32]), ("tiny_turbo_T2", True, [20, 40]),
                                           ("tiny_sd15_T4", False, [18, 26, 35, 45])])
def test_engine_matches_golden(cuda, name, turbo, tl):
    from ai_rtc_agent_b200.host import arch as A
    from ai_rtc_agent_b200.host.stream import StreamDiffusion
    from oracle import unet as ounet
    from oracle import weights as ow     # weight / frame generators only; outputs are compared with the fixture
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = ounet.tiny_config(turbo)
    emb = ow.make_prompt_embeds(cfg.cross_attention_dim)
    sd = StreamDiffusion(A.TINY_TURBO if turbo else A.TINY_SD15, ow.make_unet_weights(cfg), ow.make_taesd_weights(), tl,
                         lambda p: emb, width=128, height=128, device="cuda")
    sd.prepare("p", guidance_scale=0.0, seed=2)
    assert sd.sub_timesteps == list(gold["sub_timesteps"])
    assert np.allclose(sd.alpha_prod_t_sqrt.flatten().float().numpy(), gold["alpha"], rtol=1e-3)
    assert np.allclose(sd.c_out.flatten().float().numpy(), gold["c_out"], rtol=1e-3)
    for i in range(gold["u8"].shape[0]):
        out = sd.step_u8(ow.make_frame(128, 128, seed=i).to(cuda)).cpu().numpy()
        d = np.abs(out.astype(np.int32) - gold["u8"][i:i + 1].astype(np.int32))
        assert (d <= 2).mean() >= 0.999 and d.max() <= 8, f"{name} frame {i}: max {d.max()} frac {(d <= 2).mean():.5f}"
        eps = sd.get_tensor("eps").float().permute(0, 3, 1, 2).numpy()
        ref = gold["eps"][i]
        assert np.abs(eps - ref).max() <= 2e-2 * np.abs(ref).max(), f"{name} frame {i}: eps"
