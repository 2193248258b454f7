from ai_rtc_agent_b200.host.tracks import VideoStreamTrack  # noqa: F401
