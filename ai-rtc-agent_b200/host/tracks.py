"""Media-track adapter: the reference's `lib/tracks.py:VideoStreamTrack` (lib/tracks.py:9-38) with the one change
SURVEY.md 8f-2 asks for -- `pipeline(frame)` no longer blocks the asyncio event loop.

Same behaviour as the reference: the first `WARMUP_FRAMES` (default 10) source frames are run through the pipeline and
discarded on the first `recv()`, then `DROP_FRAMES` source frames are skipped before every processed frame, and every
`recv()` returns exactly one processed frame, in order.  Differences, both deliberate:
  * the frame is ENQUEUED (StreamDiffusionPipeline.enqueue: H2D copy, engine, D2H copy all stream-ordered) and the coroutine
    yields to the event loop until the CUDA event fires, so decoding / networking of neighbouring frames and other peers
    overlaps the GPU work instead of waiting behind a blocking call (lib/tracks.py:24,38 block the loop);
  * WARMUP_FRAMES is int-cast (in the reference a value set through the environment stays a str and `int < str` raises).
The reference's own lib/tracks.py also runs unmodified on lib.pipeline (tests/test_tracks.py); this adapter is the
non-blocking replacement.  aiortc is optional: with it installed the class derives from aiortc.MediaStreamTrack."""
from __future__ import annotations

import asyncio
import logging
import os

logger = logging.getLogger(__name__)

try:  # pragma: no cover - aiortc is not installable offline
    from aiortc import MediaStreamTrack as _Base
except ImportError:
    class _Base:  # minimal stand-in with the attributes aiortc's base class provides
        kind = "unknown"

        def __init__(self):
            self._ended = False

        def stop(self):
            self._ended = True


class VideoStreamTrack(_Base):
    kind = "video"

    def __init__(self, track, pipeline, poll_interval: float = 0.0):
        super().__init__()
        self.track = track
        self.pipeline = pipeline
        self.warmup_frame_idx = 0
        self.warmup_frames = int(os.getenv("WARMUP_FRAMES", 10))
        self.drop_frames = int(os.getenv("DROP_FRAMES", 0))
        self.poll_interval = poll_interval

    async def _process(self, frame):
        enqueue = getattr(self.pipeline, "enqueue", None)
        if enqueue is None:                      # any callable pipeline works; it is then called synchronously like the reference
            return self.pipeline(frame)
        ticket = enqueue(frame)
        while not ticket.done():
            await asyncio.sleep(self.poll_interval)   # let the event loop run while the GPU works
        return ticket.result()

    async def recv(self):
        while self.warmup_frame_idx < self.warmup_frames:
            logger.info(f"dropping warmup frames {self.warmup_frame_idx}")
            frame = await self.track.recv()
            await self._process(frame)
            self.warmup_frame_idx += 1

        # Frame dropping (lib/tracks.py:27-31): skipping source frames can help playback with some encoders
        for _ in range(self.drop_frames):
            await self.track.recv()

        frame = await self.track.recv()
        return await self._process(frame)
