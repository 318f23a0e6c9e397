"""EasyBytes on the device -- the decode half of USTC_lab/data/easybytes.py for the Forward / Backward modules
(SURVEY 8f, row f2).

The reference decodes a Redis payload on the host: ``decode_forward_states`` (easybytes.py:114-139) slices every env
process's message with ``np.frombuffer``, ``np.concatenate``s each state slot over the processes, and the Forward thread
converts every slot with ``torch.tensor(state, dtype=float32).to(device)`` (server/forward.py:128-131).  Here the host
reads only the headers (``parse_*``: pure Python, testable without a GPU); the raw payload crosses PCIe once, in its
wire dtypes (uint8 frames are 4x, float64 Pong observations 0.5x the fp32 bytes), and ONE kernel
(``ddrl_easybytes_decode``) does slice + concatenate + conversion into fp32 device tensors.

Wire format (easybytes.py:47-61,140-148), big-endian headers, little-endian data:
    data block   [type >h: 1 u8 | 2 f16 | 3 f32 | 4 f64][count >I][ndim >I][shape >I x ndim][count * size bytes]
    forward msg  [length >Q][ip 4 x >H][process_env_id >I][data blocks, `length` bytes]
    backward msg [len >Q][state blocks][len >Q][advs, actions, old_logps, values blocks][marshal(logger dict)]
The encode half (``encode_forward_return_data`` etc.) is tiny host work and stays with the reference class."""
import marshal
import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import check, current_stream, ptr

TYPE_SIZE = {1: 1, 2: 2, 3: 4, 4: 8}                     # easybytes.py:21-26
SEG_DTYPE = np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("count", "<u4"), ("dtype", "<u4")])   # = csrc Seg (24 bytes)


def parse_data(buf, start: int, end: int) -> List[Tuple[int, Tuple[int, ...], int, int]]:
    """Headers of the data blocks in buf[start:end] -> [(type code, shape, data byte offset, count)]."""
    out, i = [], start
    while i < end:
        if i + 10 > end:
            raise ValueError("EasyBytes: truncated block header at byte %d" % i)
        code = struct.unpack_from(">h", buf, i)[0]
        if code not in TYPE_SIZE:
            raise ValueError("EasyBytes: unknown type code %r at byte %d" % (code, i))      # reference: KeyError
        count, ndim = struct.unpack_from(">II", buf, i + 2)
        if ndim > 8 or i + 10 + 4 * ndim > end:
            raise ValueError("EasyBytes: bad rank %d at byte %d" % (ndim, i))
        shape = struct.unpack_from(">" + "I" * ndim, buf, i + 10)
        i += 10 + 4 * ndim
        if int(np.prod(shape, dtype=np.int64)) != count or i + count * TYPE_SIZE[code] > end:
            raise ValueError("EasyBytes: truncated or inconsistent block at byte %d" % i)
        out.append((code, tuple(shape), i, count))
        i += count * TYPE_SIZE[code]
    return out


def parse_forward_states(buf) -> Tuple[List[str], List[List[Tuple[int, Tuple[int, ...], int, int]]]]:
    """-> (process_env_ids, per message its block list)   (easybytes.py:114-130)."""
    ids, msgs, i, n = [], [], 0, len(buf)
    while i < n:
        if i + 20 > n:
            raise ValueError("EasyBytes: truncated message header at byte %d" % i)
        length = struct.unpack_from(">Q", buf, i)[0]
        if i + 20 + length > n:
            raise ValueError("EasyBytes: message at byte %d claims %d payload bytes, %d left" % (i, length, n - i - 20))
        ip = struct.unpack_from(">HHHH", buf, i + 8)
        env_id = struct.unpack_from(">I", buf, i + 16)[0]
        msgs.append(parse_data(buf, i + 20, i + 20 + length))
        ids.append(".".join(str(x) for x in ip) + "_" + str(env_id))
        i += 20 + length
    return ids, msgs


def concat_plan(msgs: Sequence[Sequence[Tuple[int, Tuple[int, ...], int, int]]], axis1_slots: Sequence[int] = ()):
    """np.concatenate of slot k over the messages (axis 0; axis 1 for the 2-D slots listed in `axis1_slots`, i.e. the
    [V, B] value rows of Experience.batch_data, data/experience.py:116-148), as a segment table:
    -> (slot shapes, slot float offsets into one flat fp32 buffer, total floats, segments structured array)."""
    n_slots = len(msgs[0])
    shapes, offs, segs, total = [], [], [], 0
    for k in range(n_slots):
        if any(len(m) != n_slots for m in msgs):
            raise ValueError("EasyBytes: messages carry different numbers of arrays")
        offs.append(total)
        if k in axis1_slots:
            V = msgs[0][k][1][0]
            if any(len(m[k][1]) != 2 or m[k][1][0] != V for m in msgs):
                raise ValueError("EasyBytes: slot %d is not [V, B] with the same V in every message" % k)
            cols = sum(m[k][1][1] for m in msgs)
            col0 = 0
            for m in msgs:
                code, shape, off, count = m[k]
                b = shape[1]
                for v in range(V):                                   # row v of this message -> dst[v, col0 : col0 + b]
                    segs.append((off + v * b * TYPE_SIZE[code], total + v * cols + col0, b, code))
                col0 += b
            shapes.append((V, cols))
            total += V * cols
            continue
        tail = msgs[0][k][1][1:]
        rows = 0
        for m in msgs:
            if m[k][1][1:] != tail:
                raise ValueError("EasyBytes: state slot %d differs between env processes" % k)   # np.concatenate would raise
            code, shape, off, count = m[k]
            segs.append((off, total, count, code))
            total += count
            rows += shape[0]
        shapes.append((rows,) + tail)
    return shapes, offs, total, np.array(segs, dtype=SEG_DTYPE)


def _copy_bytes(dst, src, b0, q0, q1):
    dst[q0:q1] = src[b0 + q0:b0 + q1]


class DeviceEasyBytes:
    """decode_forward_states / decode_backward_data with fp32 DEVICE tensors as the result."""

    N_SLOTS = 3                                                    # pinned staging ring (one event per slot)

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self._pin = [None] * self.N_SLOTS
        self._evt = [None] * self.N_SLOTS
        self._slot = 0

    def _upload(self, bufs, bases, segs: np.ndarray):
        """Stage payloads + segment table in the next pinned slot and queue ONE H2D copy.  The copy is asynchronous (it may
        sit behind training kernels when the learner prefetches batch k+1): the slot's event is recorded behind it and
        waited on before the host writes that slot again, so a queued copy never reads bytes of a later call."""
        n = bases[-1]
        need = n + segs.nbytes + 64
        k = self._slot
        self._slot = (k + 1) % self.N_SLOTS
        if self._evt[k] is not None:
            self._evt[k].synchronize()
        if self._pin[k] is None or self._pin[k].numel() < need:
            self._pin[k] = torch.empty(max(need, 1 << 20), dtype=torch.uint8).pin_memory()
        seg_off = (n + 31) // 32 * 32                              # 8-byte aligned records behind the payloads
        host = self._pin[k].numpy()
        for b, o in zip(bufs, bases):
            host[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
        host[seg_off:seg_off + segs.nbytes] = segs.view(np.uint8).reshape(-1)
        dev = self._pin[k][:seg_off + segs.nbytes].to(self.device, non_blocking=True)       # ONE H2D copy
        if self.device.type == "cuda":
            if self._evt[k] is None:
                self._evt[k] = torch.cuda.Event()
            self._evt[k].record(torch.cuda.current_stream(self.device))
        return dev, seg_off

    def _decode(self, buf, msgs, axis1_slots=(), bases=None):
        """buf: one payload, or a list of payloads with `bases` = their byte offsets in the staged buffer (+ total) --
        the block offsets inside `msgs` are then already absolute."""
        shapes, offs, total, segs = concat_plan(msgs, axis1_slots)
        lib = _lib.load()
        bufs = [buf] if bases is None else buf
        dev, seg_off = self._upload(bufs, [0, len(buf)] if bases is None else bases, segs)
        out = torch.empty(total, dtype=torch.float32, device=self.device)
        check(lib.ddrl_easybytes_decode(ptr(dev), dev.data_ptr() + seg_off, len(segs), int(segs["count"].max(initial=0)),
                                        ptr(out), current_stream()), "ddrl_easybytes_decode")
        return [out[o:o + int(np.prod(s, dtype=np.int64))].view(*s) for o, s in zip(offs, shapes)]

    def decode_forward_states(self, byte_states) -> Tuple[List[str], List[torch.Tensor]]:
        """easybytes.py:114-139 + server/forward.py:128-131: (process_env_ids, fp32 device state slots)."""
        ids, msgs = parse_forward_states(byte_states)
        return ids, self._decode(byte_states, msgs)

    # ---- large payloads: chunks of whole messages, staged copy / H2D / decode of consecutive chunks overlapped ---------
    def forward_chunks(self, byte_states, chunk_bytes: int = 64 << 20):
        """Splits a Forward payload into runs of whole env-process messages of <= chunk_bytes each.
        -> (process_env_ids, [(byte0, byte1, msgs with offsets relative to byte0)])."""
        ids, msgs, bounds, i, n = [], [], [], 0, len(byte_states)
        view = byte_states if isinstance(byte_states, (bytes, bytearray, memoryview)) else memoryview(byte_states.numpy())
        while i < n:
            if i + 20 > n:
                raise ValueError("EasyBytes: truncated message header at byte %d" % i)
            length = struct.unpack_from(">Q", view, i)[0]
            if i + 20 + length > n:
                raise ValueError("EasyBytes: message at byte %d claims %d payload bytes, %d left" % (i, length, n - i - 20))
            ip = struct.unpack_from(">HHHH", view, i + 8)
            env_id = struct.unpack_from(">I", view, i + 16)[0]
            msgs.append(parse_data(view, i + 20, i + 20 + length))
            ids.append(".".join(str(x) for x in ip) + "_" + str(env_id))
            bounds.append((i, i + 20 + length))
            i += 20 + length
        chunks, k = [], 0
        while k < len(msgs):
            b0, j = bounds[k][0], k
            while j < len(msgs) and (j == k or bounds[j][1] - b0 <= chunk_bytes):
                j += 1
            b1 = bounds[j - 1][1]
            chunks.append((b0, b1, [[(c, sh, off - b0, cnt) for c, sh, off, cnt in m] for m in msgs[k:j]]))
            k = j
        return ids, chunks

    def upload_chunk(self, src, b0: int, b1: int, segs: np.ndarray, stream, pool=None, threads: int = 8):
        """Queues the H2D copy of src[b0:b1] (+ the segment table) on `stream`.  src: a PINNED uint8 tensor (copied straight
        from where it lies) or a bytes-like object (staged through the pinned ring; `pool` splits the host copy over
        `threads` workers -- numpy releases the GIL).  -> (device payload, device segment table)."""
        n = b1 - b0
        if torch.is_tensor(src) and src.is_pinned():
            with torch.cuda.stream(stream):
                dev = src[b0:b1].to(self.device, non_blocking=True)
                seg_dev = torch.from_numpy(segs.view(np.uint8).reshape(-1).copy()).to(self.device, non_blocking=True)
            return dev, seg_dev
        k = self._slot
        self._slot = (k + 1) % self.N_SLOTS
        if self._evt[k] is not None:
            self._evt[k].synchronize()
        need = n + segs.nbytes + 64
        if self._pin[k] is None or self._pin[k].numel() < need:
            self._pin[k] = torch.empty(max(need, 1 << 20), dtype=torch.uint8).pin_memory()
        host = self._pin[k].numpy()
        raw = np.frombuffer(src, dtype=np.uint8) if not torch.is_tensor(src) else src.numpy()
        if pool is not None and n > (8 << 20):
            per = -(-n // threads)
            jobs = [pool.submit(_copy_bytes, host, raw, b0, q, min(n, q + per)) for q in range(0, n, per)]
            for j in jobs:
                j.result()
        else:
            host[:n] = raw[b0:b1]
        seg_off = (n + 31) // 32 * 32
        host[seg_off:seg_off + segs.nbytes] = segs.view(np.uint8).reshape(-1)
        with torch.cuda.stream(stream):
            dev_all = self._pin[k][:seg_off + segs.nbytes].to(self.device, non_blocking=True)
            if self._evt[k] is None:
                self._evt[k] = torch.cuda.Event()
            self._evt[k].record(stream)
        return dev_all[:n], dev_all[seg_off:]

    def decode_uploaded(self, dev, seg_dev, msgs):
        """Runs the decode kernel on an uploaded chunk (current stream).  -> fp32 device state slots of the chunk."""
        shapes, offs, total, segs = concat_plan(msgs)
        out = torch.empty(total, dtype=torch.float32, device=self.device)
        check(_lib.load().ddrl_easybytes_decode(ptr(dev), ptr(seg_dev), len(segs), int(segs["count"].max(initial=0)), ptr(out),
                                                current_stream()), "ddrl_easybytes_decode")
        return [out[o:o + int(np.prod(s, dtype=np.int64))].view(*s) for o, s in zip(offs, shapes)]

    def decode_backward_data(self, bytes_data) -> Tuple[List[torch.Tensor], List[torch.Tensor], Dict]:
        """easybytes.py:163-171: (states, [advs, actions, old_logps, values], logger dict), tensors fp32 on the device."""
        n0 = struct.unpack_from(">Q", bytes_data, 0)[0]
        st = parse_data(bytes_data, 8, 8 + n0)
        n1 = struct.unpack_from(">Q", bytes_data, 8 + n0)[0]
        other = parse_data(bytes_data, 16 + n0, 16 + n0 + n1)
        # one "message" whose slots are all blocks: no concatenation, only slicing + conversion
        tensors = self._decode(bytes_data, [st + other])
        logger = marshal.loads(bytes(bytes_data[16 + n0 + n1:]))
        return tensors[:len(st)], tensors[len(st):], logger

    def decode_backward_batch(self, payloads: Sequence[bytes]):
        """BackwardQueue.get + Experience.batch_data on the device (server/backward.py:48-62, data/experience.py:116-148):
        several training payloads -> ONE staged copy -> ONE kernel -> (Experience of concatenated fp32 device tensors,
        [logger dicts]).  States / advs / actions / old_logps concatenate along axis 0, values [V, B] along axis 1."""
        from .experience import Experience
        msgs, loggers, bases, base = [], [], [], 0
        n_states = None
        for p in payloads:
            n0 = struct.unpack_from(">Q", p, 0)[0]
            st = parse_data(p, 8, 8 + n0)
            n1 = struct.unpack_from(">Q", p, 8 + n0)[0]
            other = parse_data(p, 16 + n0, 16 + n0 + n1)
            if len(other) != 4 or (n_states is not None and len(st) != n_states):
                raise ValueError("EasyBytes: training payload does not hold [states..., advs, actions, old_logps, values]")
            n_states = len(st)
            msgs.append([(c, sh, off + base, cnt) for c, sh, off, cnt in st + other])
            loggers.append(marshal.loads(bytes(p[16 + n0 + n1:])))
            bases.append(base)
            base += (len(p) + 15) // 16 * 16
        bases.append(base)
        t = self._decode(list(payloads), msgs, axis1_slots=(n_states + 3,), bases=bases)
        exp = Experience(states=t[:n_states], advs=t[n_states], actions=t[n_states + 1], old_logps=t[n_states + 2],
                         values=t[n_states + 3])
        return exp, loggers
