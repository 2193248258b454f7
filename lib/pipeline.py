from ai_rtc_agent_b200.host.pipeline import (  # noqa: F401
    DEFAULT_GUIDANCE_SCALE, DEFAULT_NUM_INFERENCE_STEPS, DEFAULT_PROMPT, DEFAULT_T_INDEX_LIST,
    StreamDiffusionPipeline)
