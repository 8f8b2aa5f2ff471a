"""Token streams of in-flight requests (reference ``core/inference/async_stream.py`` ``AsyncStream``).

A ``TokenStream`` is handed to the engine with a request; the engine ``put``s every newly sampled token (``{"token", "text",
"logprob"}``) from whatever thread steps the model and ``finish``es the stream with the completed ``InferenceRequest``.
Consumers read it either synchronously (``for chunk in stream``: the HTTP server's SSE writer thread) or from asyncio
(``async for chunk in stream.generator()``); abandoning the iterator calls ``cancel(request_id)`` so the engine can drop
the request."""
from __future__ import annotations

import asyncio
import queue
from typing import Any, AsyncGenerator, Callable, Iterator, Optional

_END = object()


class TokenStream:
    def __init__(self, request_id: Optional[str] = None, cancel: Optional[Callable[[str], None]] = None):
        self.request_id = request_id
        self._cancel = cancel
        self._q: "queue.Queue[Any]" = queue.Queue()
        self._finished = False
        self.result = None                        # the completed InferenceRequest (set by finish)
        self._loop: Optional[asyncio.AbstractEventLoop] = None
        self._aq: Optional[asyncio.Queue] = None

    # ---- producer side (engine thread) -------------------------------------------------------------------------------
    def put(self, item: Any) -> None:
        if self._finished:
            return
        self._q.put(item)
        if self._aq is not None:
            self._loop.call_soon_threadsafe(self._aq.put_nowait, item)

    def finish(self, result: Any = None, exception: Optional[BaseException] = None) -> None:
        if self._finished:
            return
        self._finished = True
        self.result = result
        end = exception if exception is not None else _END
        self._q.put(end)
        if self._aq is not None:
            self._loop.call_soon_threadsafe(self._aq.put_nowait, end)

    @property
    def finished(self) -> bool:
        return self._finished

    # ---- consumers ---------------------------------------------------------------------------------------------------
    def __iter__(self) -> Iterator[Any]:
        try:
            while True:
                item = self._q.get()
                if item is _END:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        except GeneratorExit:
            self.cancel()
            raise

    def bind_loop(self) -> "TokenStream":
        """Call from the consuming coroutine BEFORE the request is submitted so no chunk is missed."""
        self._loop = asyncio.get_running_loop()
        self._aq = asyncio.Queue()
        return self

    async def generator(self) -> AsyncGenerator[Any, None]:
        if self._aq is None:
            self.bind_loop()
            while not self._q.empty():                       # chunks produced before the loop was bound
                self._aq.put_nowait(self._q.get_nowait())
        try:
            while True:
                item = await self._aq.get()
                if item is _END:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        except GeneratorExit:
            self.cancel()
            raise asyncio.CancelledError from None

    def cancel(self) -> None:
        if self._cancel is not None and not self._finished:
            self._cancel(self.request_id)


AsyncStream = TokenStream      # the reference's name
