"""One background asyncio loop shared by every part of the API (pylzy/lzy/utils/event_loop.py:8-44):
``Runtime`` methods are coroutines and always run on this single "workflow-thread"."""
from __future__ import annotations

import asyncio
import contextvars
import sys
from threading import Lock, Thread
from typing import Any, Awaitable, Dict, Optional, Sequence, TypeVar

T = TypeVar("T")

# CUDA streams are per thread.  A caller hands tensors in (put_data) and takes tensors out (get_data) with its own
# current stream in mind — the H2D copy that fills an argument, the D2H copy that reads a result — but the
# coroutine that touches them runs on the workflow thread.  run_async therefore carries the caller's current
# stream of every device into the coroutine's context; the frame store packs / clones on that stream
# (`caller_stream`), which keeps both sides of the hand-over in one stream order.
_CALLER_STREAMS: contextvars.ContextVar = contextvars.ContextVar("lzyb_caller_streams", default=None)


def _capture_streams() -> Optional[Dict[int, Any]]:
    torch = sys.modules.get("torch")
    if torch is None or not torch.cuda.is_available() or not torch.cuda.is_initialized():
        return None
    return {d: torch.cuda.current_stream(d) for d in range(torch.cuda.device_count())}


_INLINE = object()  # _CALLER_STREAMS while a coroutine is driven on its caller's own thread: "current" already is the caller's


def caller_stream(gpu: int) -> Optional[Any]:
    """the stream that was current on `gpu` in the thread that called run_async (None outside such a call, and None
    when the coroutine runs on the caller's thread itself: its current stream is the answer)"""
    streams = _CALLER_STREAMS.get()
    return None if streams is None or streams is _INLINE else streams.get(gpu)


class LzyEventLoop:
    instance: Optional["LzyEventLoop"] = None
    _lock = Lock()

    def __init__(self):
        self._loop = asyncio.new_event_loop()
        self._thread = Thread(name="workflow-thread", target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        asyncio.set_event_loop(self._loop)
        self._loop.run_forever()

    @classmethod
    def _get(cls) -> "LzyEventLoop":
        with cls._lock:
            if cls.instance is None:
                cls.instance = cls()
            return cls.instance

    @classmethod
    def run_async(cls, fun: Awaitable[T], inline: bool = False) -> T:
        """`inline`: the callee has declared that this coroutine (and everything it awaits) never really suspends —
        ``B200Runtime`` and ``FrameSnapshot`` do (`inline_coroutines`): their bodies are synchronous calls into the C
        ABI.  It is then driven to completion right here, on the caller's thread: no hand-off to the workflow thread and
        back (two thread switches and a self-pipe wake-up, ≈ 80 µs a time, five times per two-op workflow), no capture of
        the caller's streams (the caller's thread is the one running).  Any other runtime / snapshot — the reference's
        contract is five coroutines on one loop thread — takes the loop as before."""
        if inline:
            token = _CALLER_STREAMS.set(_INLINE)
            try:
                try:
                    fun.send(None)  # type: ignore[attr-defined]
                except StopIteration as done:
                    return done.value
                fun.close()  # type: ignore[attr-defined]
                raise RuntimeError("a coroutine declared inline-safe suspended; it needs the workflow thread")
            finally:
                _CALLER_STREAMS.reset(token)
        streams = _capture_streams()
        if streams is None:
            return asyncio.run_coroutine_threadsafe(fun, cls._get()._loop).result()

        async def with_caller_streams():
            _CALLER_STREAMS.set(streams)  # the task's own context: child tasks inherit it, other tasks do not see it
            return await fun

        return asyncio.run_coroutine_threadsafe(with_caller_streams(), cls._get()._loop).result()

    @classmethod
    def gather(cls, *args: Awaitable[T], inline: bool = False) -> Sequence[T]:
        if inline:
            return [cls.run_async(a, inline=True) for a in args]

        async def _g():
            return await asyncio.gather(*args)
        return cls.run_async(_g())
