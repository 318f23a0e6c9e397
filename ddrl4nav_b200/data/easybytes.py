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


def concat_plan(msgs: Sequence[Sequence[Tuple[int, Tuple[int, ...], int, int]]]):
    """np.concatenate(axis=0) of slot k over the messages, as a segment table:
    -> (slot shapes, slot float offsets into one flat fp32 buffer, total floats, segments structured array)."""
    n_slots = len(msgs[0])
    shapes, offs, segs, total = [], [], [], 0
    for k in range(n_slots):
        tail = msgs[0][k][1][1:]
        rows = 0
        offs.append(total)
        for m in msgs:
            if len(m) != n_slots or m[k][1][1:] != tail:
                raise ValueError("EasyBytes: state slot %d differs between env processes" % k)   # np.concatenate would raise
            code, shape, off, count = m[k]
            segs.append((off, total, count, code))
            total += count
            rows += shape[0]
        shapes.append((rows,) + tail)
    return shapes, offs, total, np.array(segs, dtype=SEG_DTYPE)


class DeviceEasyBytes:
    """decode_forward_states / decode_backward_data with fp32 DEVICE tensors as the result."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self._pin = None

    def _upload(self, buf, segs: np.ndarray):
        n = len(buf)
        need = n + segs.nbytes + 64
        if self._pin is None or self._pin.numel() < need:
            self._pin = torch.empty(max(need, 1 << 20), dtype=torch.uint8).pin_memory()
        seg_off = (n + 31) // 32 * 32                              # 8-byte aligned records behind the payload
        host = self._pin.numpy()
        host[:n] = np.frombuffer(buf, dtype=np.uint8)
        host[seg_off:seg_off + segs.nbytes] = segs.view(np.uint8).reshape(-1)
        dev = self._pin[:seg_off + segs.nbytes].to(self.device, non_blocking=True)          # ONE H2D copy
        return dev, seg_off

    def _decode(self, buf, msgs):
        shapes, offs, total, segs = concat_plan(msgs)
        lib = _lib.load()
        dev, seg_off = self._upload(buf, segs)
        out = torch.empty(total, dtype=torch.float32, device=self.device)
        check(lib.ddrl_easybytes_decode(ptr(dev), dev.data_ptr() + seg_off, len(segs), int(segs["count"].max(initial=0)),
                                        ptr(out), current_stream()), "ddrl_easybytes_decode")
        return [out[o:o + int(np.prod(s, dtype=np.int64))].view(*s) for o, s in zip(offs, shapes)]

    def decode_forward_states(self, byte_states) -> Tuple[List[str], List[torch.Tensor]]:
        """easybytes.py:114-139 + server/forward.py:128-131: (process_env_ids, fp32 device state slots)."""
        ids, msgs = parse_forward_states(byte_states)
        return ids, self._decode(byte_states, msgs)

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
