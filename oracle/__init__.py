"""CPU oracle for the per-frame img2img hot path of yondonfu/ai-rtc-agent.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this package -- as the checker or as the timed CPU
baseline, never as part of the shipped path (the product path is libb200sd.so and raises if it is
missing).

PARITY UNPINNED.  The reference repository holds no tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8c), and the arithmetic lives in dependencies that are neither vendored under
/root/reference nor installable here:
  * streamdiffusion  -- git+https://github.com/yondonfu/StreamDiffusion.git@deepstream
                        (requirements.txt:14; a branch ref, no commit pin)
  * diffusers        -- pulled transitively (upstream StreamDiffusion pins diffusers==0.24.0)
  * TensorRT engines -- lib/wrapper.py:409-512
So this package restates the *published algorithms* of those versions in plain fp32 PyTorch
(UNet2DConditionModel, AutoencoderTiny, LCMScheduler tables, StreamDiffusion's stream-batch loop,
VaeImageProcessor) and anchors them on the reference's own call sites (lib/pipeline.py:50-96,
lib/wrapper.py:133-407).  What *can* be pinned without the reference is pinned in
tests/test_oracle.py: parameter counts of the two UNets (859.52 M / 865.91 M), the LCM timestep
table ([18,26,35,45] -> [639,479,299,99]), closed-form scheduler constants, and cross-checks of each
module against independent torch.nn building blocks.
"""
