"""EasyBytes wire format: oracle restatement and host-side header parser against bytes produced by the
reference's own encoder (tests/golden/easybytes.npz, made by tests/golden/make_golden.py), and the device decoder
(one H2D copy + ddrl_easybytes_decode) against the same, bit for bit."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import restate as R
from ddrl4nav_b200.data import easybytes as E


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "easybytes.npz"))


def test_oracle_decodes_reference_bytes(g):
    ids, slots = R.easybytes_decode_forward_states(g["fwd_bytes"].tobytes())
    assert ids == [str(x) for x in g["fwd_ids"]]
    for i, s in enumerate(slots):
        ref = g["fwd_slot_%d" % i]
        assert s.dtype == ref.dtype and s.shape == ref.shape and np.array_equal(s, ref)
    _, f32 = R.easybytes_forward_states_fp32(g["fwd_bytes"].tobytes())
    for i, s in enumerate(f32):
        assert s.dtype == np.float32 and np.array_equal(s, g["fwd_slot_%d_f32" % i])
    st, other, logd = R.easybytes_decode_backward_data(g["bwd_bytes"].tobytes())
    assert all(np.array_equal(a, g["bwd_state_%d" % i]) for i, a in enumerate(st))
    assert all(np.array_equal(a, g["bwd_other_%d" % i]) for i, a in enumerate(other)) and len(other) == 4
    assert sorted(logd) == [str(x) for x in g["bwd_logger_keys"]]


def test_host_parser_matches_oracle(g):
    buf = g["fwd_bytes"].tobytes()
    ids, msgs = E.parse_forward_states(buf)
    assert ids == [str(x) for x in g["fwd_ids"]] and len(msgs) == 3
    shapes, offs, total, segs = E.concat_plan(msgs)
    assert shapes == [tuple(g["fwd_slot_%d" % i].shape) for i in range(4)]
    assert total == sum(int(np.prod(s)) for s in shapes) and len(segs) == 12 and segs.dtype.itemsize == 24
    # a CPU replay of the segment table = what the kernel does
    flat = np.empty(total, np.float32)
    raw = np.frombuffer(buf, np.uint8)
    for s in segs:
        dt = R.EASYBYTES_TYPES[int(s["dtype"])][1]
        n, off = int(s["count"]), int(s["src_off"])
        flat[int(s["dst_off"]):int(s["dst_off"]) + n] = np.frombuffer(raw[off:off + n * np.dtype(dt).itemsize].tobytes(), dt).astype(np.float32)
    for i, (o, sh) in enumerate(zip(offs, shapes)):
        assert np.array_equal(flat[o:o + int(np.prod(sh))].reshape(sh), g["fwd_slot_%d_f32" % i])


def test_parser_error_behaviour(g):
    buf = bytearray(g["fwd_bytes"].tobytes())
    with pytest.raises(ValueError):
        E.parse_forward_states(bytes(buf[:-3]))                       # truncated payload
    bad = bytearray(buf)
    struct.pack_into(">h", bad, 20, 9)                                # unknown dtype code (the reference raises KeyError)
    with pytest.raises(ValueError):
        E.parse_forward_states(bytes(bad))
    ids, msgs = E.parse_forward_states(bytes(buf))
    msgs[1] = msgs[1][:-1]                                            # one env process sends fewer slots
    with pytest.raises(ValueError):
        E.concat_plan(msgs)
    assert E.parse_forward_states(b"") == ([], [])


@pytest.mark.gpu
def test_device_decode_forward_states_bit_exact(g):
    dec = E.DeviceEasyBytes("cuda:0")
    for _ in range(2):                                                # second call re-uses the pinned staging buffer
        ids, slots = dec.decode_forward_states(g["fwd_bytes"].tobytes())
        assert ids == [str(x) for x in g["fwd_ids"]]
        for i, s in enumerate(slots):
            assert s.dtype == torch.float32 and s.is_cuda
            assert np.array_equal(s.cpu().numpy(), g["fwd_slot_%d_f32" % i])


@pytest.mark.gpu
def test_device_decode_backward_data_bit_exact(g):
    dec = E.DeviceEasyBytes("cuda:0")
    st, other, logd = dec.decode_backward_data(g["bwd_bytes"].tobytes())
    for i, s in enumerate(st):
        assert np.array_equal(s.cpu().numpy(), g["bwd_state_%d" % i].astype(np.float32))
    for i, s in enumerate(other):
        assert np.array_equal(s.cpu().numpy(), g["bwd_other_%d" % i].astype(np.float32))
    assert sorted(logd) == [str(x) for x in g["bwd_logger_keys"]]


@pytest.mark.gpu
def test_device_decode_large_mixed_payload_matches_oracle():
    """Reference-shaped Pong tick: 64 env processes x 4 envs, float64 frames (warputils.py:300) + odd-sized side slots so
    that segments start at every byte alignment."""
    rng = np.random.default_rng(3)

    def block(a):
        code = {np.dtype(np.uint8): 1, np.dtype(np.float16): 2, np.dtype(np.float32): 3, np.dtype(np.float64): 4}[a.dtype]
        return struct.pack(">h", code) + struct.pack(">II", a.size, a.ndim) + struct.pack(">" + "I" * a.ndim, *a.shape) + a.tobytes()
    buf = b""
    for pid in range(64):
        n = 4
        arrs = [rng.random((n, 4, 84, 84)), rng.integers(0, 256, (n, 3), dtype=np.uint8),
                rng.standard_normal((n, 5)).astype(np.float16), rng.standard_normal((n, 7)).astype(np.float32)]
        body = b"".join(block(a) for a in arrs)
        buf += struct.pack(">Q", len(body)) + struct.pack(">HHHH", 127, 0, 0, 1) + struct.pack(">I", pid) + body
    ids_ref, ref = R.easybytes_forward_states_fp32(buf)
    ids, slots = E.DeviceEasyBytes("cuda:0").decode_forward_states(buf)
    assert ids == ids_ref
    for s, r in zip(slots, ref):
        assert tuple(s.shape) == r.shape and np.array_equal(s.cpu().numpy(), r)


@pytest.mark.gpu
def test_forward_module_step_bytes_equals_step_on_decoded_arrays():
    from ddrl4nav_b200.runner import make_net
    from ddrl4nav_b200.server import ForwardModule
    rng = np.random.default_rng(5)
    net = make_net("pong", device="cuda:0")
    net.load_state_dict(R.init_params(R.SPECS["pong"], seed=1))

    def block(a):
        return struct.pack(">h", 4) + struct.pack(">II", a.size, a.ndim) + struct.pack(">" + "I" * a.ndim, *a.shape) + a.tobytes()
    buf = b""
    for pid in range(5):
        body = block(rng.random((3, 4, 84, 84)))                     # float64 frames, 3 envs per process
        buf += struct.pack(">Q", len(body)) + struct.pack(">HHHH", 127, 0, 0, 1) + struct.pack(">I", pid) + body
    _, host_states = R.easybytes_decode_forward_states(buf)
    u = torch.rand(15, generator=torch.Generator().manual_seed(2)).to("cuda:0")
    fm = ForwardModule(net, device="cuda:0")
    ids, out_b = fm.step_bytes(buf, draw=u)
    out_a = fm.step(host_states, draw=u)
    assert ids == ["127.0.0.1_%d" % i for i in range(5)]
    for a, b in zip(out_a, out_b):
        assert np.array_equal(a, b)


def _enc(a):
    code = {np.dtype(np.uint8): 1, np.dtype(np.float16): 2, np.dtype(np.float32): 3, np.dtype(np.float64): 4}[a.dtype]
    return struct.pack(">h", code) + struct.pack(">II", a.size, a.ndim) + struct.pack(">" + "I" * a.ndim, *a.shape) + a.tobytes()


def _train_payload(rng, B, V=1):
    import marshal
    states = [rng.integers(0, 256, size=(B, 3, 4), dtype=np.uint8), rng.standard_normal((B, 4)).astype(np.float32)]
    other = [rng.standard_normal(B).astype(np.float32), rng.integers(0, 6, size=B).astype(np.float32),
             rng.standard_normal(B).astype(np.float32), rng.standard_normal((V, B)).astype(np.float32)]
    bs = b"".join(_enc(a) for a in states)
    bo = b"".join(_enc(a) for a in other)
    return struct.pack(">Q", len(bs)) + bs + struct.pack(">Q", len(bo)) + bo + marshal.dumps({"B": B}), states, other


def test_backward_batch_plan_matches_batch_data(g):
    """Segment plan of several training payloads (values concatenate along axis 1) replayed on the CPU =
    Experience.batch_data of the oracle-decoded payloads; the reference-encoded golden payload is one of them."""
    from ddrl4nav_b200.data import Experience
    rng = np.random.default_rng(11)
    pay = [g["bwd_bytes"].tobytes()] + [_train_payload(rng, B)[0] for B in (1, 9)]
    exps = []
    for p in pay:
        st, other, _ = R.easybytes_decode_backward_data(p)
        exps.append(Experience(states=st, advs=other[0], actions=other[1], old_logps=other[2], values=other[3]))
    ref = Experience.batch_data(exps)
    msgs, base = [], 0
    raw = bytearray()
    for p in pay:
        n0 = struct.unpack_from(">Q", p, 0)[0]
        n1 = struct.unpack_from(">Q", p, 8 + n0)[0]
        blocks = E.parse_data(p, 8, 8 + n0) + E.parse_data(p, 16 + n0, 16 + n0 + n1)
        msgs.append([(c, sh, off + base, cnt) for c, sh, off, cnt in blocks])
        raw += p
        base += len(p)
    shapes, offs, total, segs = E.concat_plan(msgs, axis1_slots=(5,))
    flat = np.empty(total, np.float32)
    rawa = np.frombuffer(bytes(raw), np.uint8)
    for s_ in segs:
        dt = R.EASYBYTES_TYPES[int(s_["dtype"])][1]
        n, off = int(s_["count"]), int(s_["src_off"])
        flat[int(s_["dst_off"]):int(s_["dst_off"]) + n] = np.frombuffer(rawa[off:off + n * np.dtype(dt).itemsize].tobytes(), dt)
    got = [flat[o:o + int(np.prod(sh))].reshape(sh) for o, sh in zip(offs, shapes)]
    want = list(ref.states) + [ref.advs, ref.actions, ref.old_logps, ref.values]
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.array_equal(a, b.astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("V", [1, 2])
def test_device_backward_batch_equals_batch_data(V):
    from ddrl4nav_b200.data import Experience
    rng = np.random.default_rng(12)
    made = [_train_payload(rng, B, V) for B in (5, 1, 300)]
    exp, loggers = E.DeviceEasyBytes("cuda:0").decode_backward_batch([m[0] for m in made])
    ref = Experience.batch_data([Experience(states=m[1], advs=m[2][0], actions=m[2][1], old_logps=m[2][2], values=m[2][3])
                                 for m in made])
    assert [d["B"] for d in loggers] == [5, 1, 300] and len(exp) == 306
    for a, b in zip(list(exp.states) + [exp.advs, exp.actions, exp.old_logps, exp.values],
                    list(ref.states) + [ref.advs, ref.actions, ref.old_logps, ref.values]):
        assert tuple(a.shape) == b.shape and np.array_equal(a.cpu().numpy(), b.astype(np.float32))
