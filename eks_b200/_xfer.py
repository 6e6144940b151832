"""Host <-> device transfers for the reference-facing entry points (NumPy arrays in, NumPy-backed DataFrames out).

The public functions receive PAGEABLE host arrays and must return pageable ones, so the copies cannot be plain
pinned-memory DMAs.  A pageable cudaMemcpy stages through a small driver buffer single-threaded (6-12 GB/s); here the
staging is explicit: two cached 64 MB pinned buffers, a multi-threaded host memcpy into / out of one of them while the
DMA engine drains / fills the other on a side stream.  Plumbing only -- no arithmetic.
"""

from __future__ import annotations

import numpy as np
import torch

CHUNK_BYTES = 64 << 20
_SMALL = 8 << 20
_state: dict = {}


def _staging(dev: torch.device):
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    st = _state.get(key)
    if st is None:
        bufs = [torch.empty(CHUNK_BYTES, dtype=torch.uint8).pin_memory() for _ in range(2)]
        st = dict(bufs=bufs, stream=torch.cuda.Stream(device=dev), events=[torch.cuda.Event() for _ in range(2)])
        _state[key] = st
    return st


def to_device(arr: np.ndarray, dev: torch.device) -> torch.Tensor:
    """Contiguous NumPy array -> device tensor of the same dtype and shape (asynchronous w.r.t. the current stream:
    the returned tensor is ready for work enqueued on the current stream after this call)."""
    arr = np.ascontiguousarray(arr)
    src = torch.from_numpy(arr)
    nbytes = arr.nbytes
    if nbytes < _SMALL:
        return src.to(dev)
    st = _staging(dev)
    dst = torch.empty(arr.shape, dtype=src.dtype, device=dev)
    sb = src.view(-1).view(torch.uint8)
    db = dst.view(-1).view(torch.uint8)
    side, bufs, evs = st['stream'], st['bufs'], st['events']
    side.wait_stream(torch.cuda.current_stream(dev))
    for i, off in enumerate(range(0, nbytes, CHUNK_BYTES)):
        n = min(CHUNK_BYTES, nbytes - off)
        b = i & 1
        if i >= 2:
            evs[b].synchronize()                       # the DMA that last read this buffer has finished
        bufs[b][:n].copy_(sb[off:off + n])              # multi-threaded host memcpy, overlaps the other buffer's DMA
        with torch.cuda.stream(side):
            db[off:off + n].copy_(bufs[b][:n], non_blocking=True)
            evs[b].record(side)
    torch.cuda.current_stream(dev).wait_stream(side)
    for e in evs:                                       # the staging buffers are reused by the next call
        e.synchronize()
    return dst


def to_host(t: torch.Tensor) -> np.ndarray:
    """Contiguous device tensor -> fresh pageable NumPy array (synchronous)."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    dev = t.device
    if nbytes < _SMALL:
        return t.cpu().numpy()
    st = _staging(dev)
    out = torch.empty(t.shape, dtype=t.dtype)
    ob = out.view(-1).view(torch.uint8)
    tb = t.view(-1).view(torch.uint8)
    side, bufs, evs = st['stream'], st['bufs'], st['events']
    side.wait_stream(torch.cuda.current_stream(dev))
    offs = list(range(0, nbytes, CHUNK_BYTES))

    def issue(i):
        n = min(CHUNK_BYTES, nbytes - offs[i])
        with torch.cuda.stream(side):
            bufs[i & 1][:n].copy_(tb[offs[i]:offs[i] + n], non_blocking=True)
            evs[i & 1].record(side)

    issue(0)
    for i in range(len(offs)):
        evs[i & 1].synchronize()                        # chunk i has landed in its pinned buffer
        if i + 1 < len(offs):
            issue(i + 1)                                # next DMA into the other buffer while this one is copied out
        n = min(CHUNK_BYTES, nbytes - offs[i])
        ob[offs[i]:offs[i] + n].copy_(bufs[i & 1][:n])  # multi-threaded host memcpy (also first-touches the pages)
    return out.numpy()
