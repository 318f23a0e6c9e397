"""Forward module compute body -- USTC_lab/server/forward.py:128-170.

``ForwardModule`` owns what the reference's ``ForwardThread.run`` does between popping a Redis item
and pushing the replies: host arrays -> pinned staging -> device, ONE fused engine call
(encoders + heads + sampling + log-prob + value), device -> host.  Redis / EasyBytes stay with the
caller (the reference thread); INTEGRATION.md shows the 6-line patch."""
from typing import List, Optional, Sequence

import numpy as np
import torch


def forward_compute(net, batch_states: Sequence, play_mode: bool = False, draw: Optional[torch.Tensor] = None):
    """(actions, logps, values [V,B,1]) as fp32 device tensors -- forward.py:132-146."""
    return net.act(batch_states, draw=draw, play_mode=play_mode)


class ForwardModule:
    """chunk_bytes: batches whose observations exceed it are streamed in row chunks through a ring of pinned staging
    buffers -- host copy of chunk c+1 (numpy releases the GIL: `copy_threads` workers fill disjoint row ranges), H2D
    DMA of chunk c on a copy stream and the engine call of chunk c-1 all overlap."""

    def __init__(self, net, play_mode: bool = False, nptype=np.float32, device="cuda", chunk_bytes: int = 96 << 20,
                 copy_threads: int = 8):
        self.net = net
        self.play_mode = play_mode
        self.nptype = nptype
        self.device = torch.device(device)
        self.chunk_bytes = int(chunk_bytes)
        self.copy_threads = max(1, int(copy_threads))
        self._pinned = {}
        self._ring = {}
        self._pool = None
        self._copy_stream = None

    def _stage(self, i: int, a) -> torch.Tensor:
        """np array (u8/f16/f32/f64, easybytes.py:21-26) -> fp32 device tensor through a reusable pinned buffer."""
        if torch.is_tensor(a):
            return a.to(self.device, torch.float32, non_blocking=True)
        a = np.asarray(a)
        key = (i, a.shape, a.dtype.str)
        buf = self._pinned.get(key)
        if buf is None:
            buf = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype).pin_memory()
            self._pinned = {k: v for k, v in self._pinned.items() if k[0] != i}
            self._pinned[key] = buf
        buf.numpy()[...] = a
        return buf.to(self.device, non_blocking=True).to(torch.float32)

    def _finish(self, actions, logps, values):
        out = torch.cat([actions.reshape(-1), logps, values.reshape(-1)]).cpu().numpy().astype(self.nptype, copy=False)
        B = logps.shape[0]
        na = actions.numel()
        return [out[:na].reshape(tuple(actions.shape)), out[na:na + B], out[na + B:].reshape(1, B, 1)]

    def step(self, batch_states: Sequence, draw: Optional[torch.Tensor] = None) -> List[np.ndarray]:
        """Returns [actions, logps, values] as numpy (the list ``encode_forward_return_data`` consumes)."""
        host = all(not torch.is_tensor(s) for s in batch_states)
        if host:
            arrs = [np.asarray(s) for s in batch_states]
            B = len(arrs[0])
            row_bytes = sum(a.nbytes // max(B, 1) for a in arrs)
            if B * row_bytes > self.chunk_bytes and B >= 512:
                return self._step_streamed(arrs, B, row_bytes, draw)
        states = [self._stage(i, s) for i, s in enumerate(batch_states)]
        actions, logps, values = self.net.act(states, draw=draw, play_mode=self.play_mode)
        return self._finish(actions, logps, values)

    def step_bytes(self, byte_states, draw: Optional[torch.Tensor] = None):
        """One tick straight from the Redis payload (server/forward.py:117-146): the wire bytes cross PCIe once in their wire
        dtypes and are sliced / concatenated / converted to fp32 by ONE kernel (data/easybytes.py), then the fused engine
        call.  Returns (process_env_ids, [actions, logps, values]) -- the two things ``ForwardThread.run`` needs for
        ``encode_forward_return_data`` and the reply keys."""
        ids, (actions, logps, values) = self._run_bytes(byte_states, draw)
        return ids, self._finish(actions, logps, values)

    def step_bytes_replies(self, byte_states, per_env: int, draw: Optional[torch.Tensor] = None):
        """``step_bytes`` + ``encode_forward_return_data`` (data/easybytes.py:77-109) with the replies cut on the DEVICE:
        one kernel writes every env process's reply (headers + fp32 payloads) into one byte buffer, ONE device->host copy
        brings them back, the host slices it.  -> (process_env_ids, [reply bytes per env process])."""
        import ctypes as C
        from .. import _lib
        from .._lib import check, current_stream, ptr
        ids, (actions, logps, values) = self._run_bytes(byte_states, draw)
        lib = _lib.load()
        n_env, B = len(ids), logps.shape[0]
        if n_env * per_env != B:
            raise ValueError("%d env processes x %d rows != batch of %d rows" % (n_env, per_env, B))
        A = actions.shape[1] if actions.dim() == 2 else 0
        rb = int(lib.ddrl_easybytes_reply_bytes(per_env, A, 1))
        out = torch.empty(n_env * rb, dtype=torch.uint8, device=self.device)
        check(lib.ddrl_easybytes_encode_replies(ptr(actions), A, ptr(logps), ptr(values), 1, B, n_env, per_env, ptr(out),
                                                current_stream()), "ddrl_easybytes_encode_replies")
        key = ("replies", out.numel())
        host = self._pinned.get(key)
        if host is None:
            host = torch.empty(out.numel(), dtype=torch.uint8).pin_memory()
            self._pinned[key] = host
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        raw = host.numpy().tobytes()
        return ids, [raw[j * rb:(j + 1) * rb] for j in range(n_env)]

    def _run_bytes(self, byte_states, draw):
        """decode + engine call for a payload; payloads above `chunk_bytes` stream through in chunks of whole messages:
        host staging (threaded) / H2D on the copy stream of chunk c+1 overlap decode + engine call of chunk c."""
        if getattr(self, "_eb", None) is None:
            from ..data.easybytes import DeviceEasyBytes
            self._eb = DeviceEasyBytes(self.device)
        n = byte_states.numel() if torch.is_tensor(byte_states) else len(byte_states)
        if n <= self.chunk_bytes:
            if torch.is_tensor(byte_states):
                byte_states = byte_states.numpy().tobytes()
            ids, states = self._eb.decode_forward_states(byte_states)
            return ids, self.net.act(states, draw=draw, play_mode=self.play_mode)
        from concurrent.futures import ThreadPoolExecutor
        from ..data.easybytes import concat_plan
        if self._pool is None:
            self._pool = ThreadPoolExecutor(self.copy_threads)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        main = torch.cuda.current_stream(self.device)
        ids, chunks = self._eb.forward_chunks(byte_states, self.chunk_bytes)
        outs, row0 = [], 0
        nxt = None
        for c, (b0, b1, msgs) in enumerate(chunks):
            if nxt is None:
                segs = concat_plan(msgs)[3]
                nxt = self._eb.upload_chunk(byte_states, b0, b1, segs, self._copy_stream, self._pool, self.copy_threads)
                ev = torch.cuda.Event(); ev.record(self._copy_stream)
                nxt = nxt + (ev,)
            dev, seg_dev, ev = nxt
            nxt = None
            if c + 1 < len(chunks):                      # start the next chunk's staging + H2D before this chunk computes
                nb0, nb1, nmsgs = chunks[c + 1]
                up = self._eb.upload_chunk(byte_states, nb0, nb1, concat_plan(nmsgs)[3], self._copy_stream, self._pool, self.copy_threads)
                nev = torch.cuda.Event(); nev.record(self._copy_stream)
                nxt = up + (nev,)
            main.wait_event(ev)
            dev.record_stream(main); seg_dev.record_stream(main)
            states = self._eb.decode_uploaded(dev, seg_dev, msgs)
            rows = states[0].shape[0]
            dr = None if draw is None else draw[row0:row0 + rows]
            outs.append(self.net.act(states, draw=dr, play_mode=self.play_mode))
            row0 += rows
        actions = torch.cat([o[0] for o in outs], 0)
        logps = torch.cat([o[1] for o in outs], 0)
        values = torch.cat([o[2] for o in outs], 1)
        return ids, (actions, logps, values)

    def _step_streamed(self, arrs, B, row_bytes, draw):
        from concurrent.futures import ThreadPoolExecutor
        rows = max(256, (self.chunk_bytes // max(row_bytes, 1)) // 128 * 128)
        if self._pool is None:
            self._pool = ThreadPoolExecutor(self.copy_threads)
            self._copy_stream = torch.cuda.Stream(self.device)
        key = (rows, tuple((a.shape[1:], a.dtype.str) for a in arrs))
        if self._ring.get("key") != key:
            self._ring = {"key": key, "slots": [
                ([torch.empty((rows,) + a.shape[1:], dtype=torch.from_numpy(a[:0]).dtype).pin_memory() for a in arrs],
                 torch.cuda.Event()) for _ in range(3)]}
        main = torch.cuda.current_stream(self.device)
        outs = []
        for c, r0 in enumerate(range(0, B, rows)):
            r1 = min(B, r0 + rows)
            n = r1 - r0
            bufs, ev = self._ring["slots"][c % 3]
            ev.synchronize()                               # the DMA that last read this slot has finished
            # host copy, split over worker threads by row range
            per = -(-n // self.copy_threads)
            jobs = [self._pool.submit(_copy_rows, a, b.numpy(), r0, q, min(n, q + per))
                    for a, b in zip(arrs, bufs) for q in range(0, n, per)]
            for j in jobs:
                j.result()
            with torch.cuda.stream(self._copy_stream):
                dev = [b[:n].to(self.device, non_blocking=True) for b in bufs]
                ev.record(self._copy_stream)
            main.wait_event(ev)
            dev32 = []
            for d in dev:
                d.record_stream(main)
                dev32.append(d if d.dtype == torch.float32 else d.to(torch.float32))
            dr = None if draw is None else draw[r0:r1]
            outs.append(self.net.act(dev32, draw=dr, play_mode=self.play_mode))
        actions = torch.cat([o[0] for o in outs], 0)
        logps = torch.cat([o[1] for o in outs], 0)
        values = torch.cat([o[2] for o in outs], 1)
        return self._finish(actions, logps, values)


def _copy_rows(src, dst, r0, q0, q1):
    dst[q0:q1] = src[r0 + q0:r0 + q1]
