from ai_rtc_agent_b200.host.wrapper import CudaStreamPtr, StreamDiffusionWrapper, postprocess_image  # noqa: F401
