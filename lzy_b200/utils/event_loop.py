"""One background asyncio loop shared by every part of the API (pylzy/lzy/utils/event_loop.py:8-44):
``Runtime`` methods are coroutines and always run on this single "workflow-thread"."""
from __future__ import annotations

import asyncio
from threading import Lock, Thread
from typing import Awaitable, Optional, Sequence, TypeVar

T = TypeVar("T")


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
    def run_async(cls, fun: Awaitable[T]) -> T:
        return asyncio.run_coroutine_threadsafe(fun, cls._get()._loop).result()

    @classmethod
    def gather(cls, *args: Awaitable[T]) -> Sequence[T]:
        async def _g():
            return await asyncio.gather(*args)
        return cls.run_async(_g())
