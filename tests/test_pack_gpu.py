"""Weight-pack CLI + packed-blob cache on the GPU (replaces build.py:11-32 and the engine cache of
lib/wrapper.py:583-615, 889-910): checkpoint on disk -> `python -m ai_rtc_agent_b200.pack` -> blob; a wrapper started from
the blob must produce bit-identical frames to one started from the checkpoint, without holding the raw weights."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _wrapper(model_dir, taesd_dir, lcm_dir, style_path, engine_dir):
    from lib.wrapper import StreamDiffusionWrapper
    w = StreamDiffusionWrapper(model_id_or_path=model_dir, t_index_list=[18, 26, 35, 45], lora_dict={style_path: 0.5},
                               lcm_lora_id=lcm_dir, vae_id=taesd_dir, width=128, height=128, output_type="pt", engine_dir=engine_dir)
    w.prepare(prompt="a prompt", num_inference_steps=50, guidance_scale=0.0)
    return w


def _frames(w, n=5):
    from oracle import weights as ow
    return [w.stream.step_u8(ow.make_frame(128, 128, seed=i).cuda()).cpu() for i in range(n)]


def test_pack_cli_blob_roundtrip(cuda, tmp_path, monkeypatch):
    from ai_rtc_agent_b200 import pack
    from tests.test_host import _write_tiny_checkpoint
    monkeypatch.setenv("B200SD_SYNTHETIC_WEIGHTS", "1")   # no text_encoder/ in the fixture: synthetic prompt embeddings
    model_dir, taesd_dir, lcm_dir, style_path, *_ = _write_tiny_checkpoint(str(tmp_path))
    engine_dir = str(tmp_path / "engines")
    # reference: weights from the checkpoint, cache disabled
    monkeypatch.setenv("B200SD_PACK_CACHE", "0")
    ref = _wrapper(model_dir, taesd_dir, lcm_dir, style_path, engine_dir)
    assert ref.packed_blob is None and not os.path.exists(engine_dir)
    want = _frames(ref)
    monkeypatch.delenv("B200SD_PACK_CACHE")
    # the CLI writes the blob ...
    rc = pack.main(["--model-id", model_dir, "--lora", f"{style_path}:0.5", "--lcm-lora-id", lcm_dir, "--vae-id", taesd_dir,
                    "--engine-dir", engine_dir, "--width", "128", "--height", "128"])
    assert rc == 0
    blobs = [os.path.join(dp, f) for dp, _, fs in os.walk(engine_dir) for f in fs]
    assert len(blobs) == 1 and blobs[0].endswith(".b2pack") and "engines--" in blobs[0]
    # ... and the next start loads it: no state dict ever reaches the engine
    from ai_rtc_agent_b200.host import weights as W
    called = []
    monkeypatch.setattr(W, "load_unet", lambda *a, **k: called.append("unet") or (_ for _ in ()).throw(AssertionError("checkpoint was read")))
    got = _wrapper(model_dir, taesd_dir, lcm_dir, style_path, engine_dir)
    assert got.packed_blob == blobs[0] and not called
    for a, b in zip(_frames(got), want):
        assert torch.equal(a, b), "frames from the packed blob must be bit-identical to frames from the checkpoint"
    # a different LoRA scale is a different recipe: must not hit this blob
    other = W.packed_blob_path(engine_dir, model_dir, "tiny-sd15", True, lcm_dir, {style_path: 0.7}, taesd_dir, False)
    assert other != blobs[0]
    # a corrupt blob falls back to the checkpoint (lib/wrapper.py:611-615 policy)
    monkeypatch.undo()
    monkeypatch.setenv("B200SD_SYNTHETIC_WEIGHTS", "1")
    with open(blobs[0], "r+b") as f:
        f.truncate(os.path.getsize(blobs[0]) // 2)
    again = _wrapper(model_dir, taesd_dir, lcm_dir, style_path, engine_dir)
    for a, b in zip(_frames(again), want):
        assert torch.equal(a, b)


def test_raw_weights_are_released_after_prepare(cuda):
    """The pack-only raw parameters (3x3 convs, q/k/v, GEGLU) are dropped after the first prepare: a second prepare still
    works (packed caches), loading further tensors is refused."""
    import ctypes as C
    from ai_rtc_agent_b200.host import arch as A
    from ai_rtc_agent_b200.host import capi
    from ai_rtc_agent_b200.host.stream import StreamDiffusion
    from oracle import unet as ounet
    from oracle import weights as ow
    cfg = ounet.tiny_config(True)
    usd, vsd, emb = ow.make_unet_weights(cfg), ow.make_taesd_weights(), ow.make_prompt_embeds(cfg.cross_attention_dim)
    sd = StreamDiffusion(A.TINY_TURBO, usd, vsd, [32], lambda p: emb, width=128, height=128)
    sd.prepare("p", guidance_scale=0.0)
    f = ow.make_frame(128, 128, seed=1).cuda()
    a = sd.step_u8(f).cpu()
    sd.prepare("p", guidance_scale=0.0)       # rebuilds the frame program from the packed caches
    assert torch.equal(sd.step_u8(f).cpu(), a)
    t = usd["conv_in.weight"]
    shape = (C.c_int64 * 4)(*t.shape)
    rc = sd._lib.b2sd_load_tensor(sd._handle, b"conv_in.weight", t.data_ptr(), 0, shape, 4)
    assert rc != 0 and b"released" in capi.lib().b2sd_last_error()


def test_checkpoint_with_real_text_encoder_end_to_end(cuda, tmp_path, monkeypatch):
    """No synthetic shortcut anywhere: UNet / TAESD / LoRAs from safetensors on disk, prompts through the checkpoint's own CLIP
    text encoder (lib/wrapper.py:468-473) on prepare and on update_prompt (lib/pipeline.py:44-45, agent.py:166-168); frames
    must track the oracle fed with the same embeddings."""
    from ai_rtc_agent_b200.host import arch as A
    from ai_rtc_agent_b200.host import weights as W
    from ai_rtc_agent_b200.host.prompt import ClipPromptEncoder
    from oracle import pipeline as opipe
    from oracle import stream as ostream
    from oracle import unet as ounet
    from oracle import weights as ow
    from tests.test_host import _write_tiny_checkpoint, _write_tiny_clip
    monkeypatch.delenv("B200SD_SYNTHETIC_WEIGHTS", raising=False)
    model_dir, taesd_dir, lcm_dir, style_path, *_ = _write_tiny_checkpoint(str(tmp_path))
    _write_tiny_clip(model_dir, hidden=A.TINY_SD15.cross_attention_dim)
    w = _wrapper(model_dir, taesd_dir, lcm_dir, style_path, str(tmp_path / "engines"))
    assert isinstance(w.stream.prompt_encoder, ClipPromptEncoder)
    # the oracle on the same (LoRA-fused) weights and the encoder's own embeddings
    _, usd, vsd, _ = W.resolve_weights(model_dir, taesd_dir, lcm_dir, True, {style_path: 0.5}, sd_turbo=False)
    cfg = ounet.tiny_config(False)
    orc = ostream.StreamOracle(ow.to_float(usd), cfg, ow.to_float(vsd), [18, 26, 35, 45], 128, 128)
    orc.prepare(w.stream.prompt_embeds[:1].float().cpu(), guidance_scale=0.0, init_noise=w.stream.init_noise.float())

    def check(i):
        f = ow.make_frame(128, 128, seed=i)
        d = (w.stream.step_u8(f.cuda()).cpu().int() - opipe.frame_to_u8(orc, f).int()).abs()
        assert (d <= 2).float().mean().item() >= 0.999 and d.max().item() <= 8, (i, d.max().item())

    for i in range(3):
        check(i)
    before = w.stream.prompt_embeds.clone()
    w.stream.update_prompt("a watercolor painting of a harbour")
    assert not torch.equal(before, w.stream.prompt_embeds)
    orc.update_prompt_embeds(w.stream.prompt_embeds[:1].float().cpu())
    for i in range(3, 6):
        check(i)
