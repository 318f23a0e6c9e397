"""Thread-level drop-ins (ddrl4nav_b200/server/threads.py) for ForwardThread / BackwardGetDataThread / BackwardTrainThread
(USTC_lab/server/forward.py:20-29,107-183, server/backward.py:68-76,141-217).

CPU: the reply encoder and the oracle's encoder restatements against bytes produced by the reference's own EasyBytes
(tests/golden/easybytes.npz).  GPU: one tick of each thread against an in-memory Redis, checked against the oracle."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import restate as R
from oracle.ref_shim import _FakeRedis


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "easybytes.npz"))


def test_reply_encoder_matches_reference_bytes(g):
    from ddrl4nav_b200.server.threads import encode_forward_replies
    for tag in ("cat", "gauss"):
        arrs = [g["reply_%s_in_%d" % (tag, k)] for k in range(3)]
        ours = encode_forward_replies(arrs, [2, 1, 3])
        orc = R.easybytes_encode_forward_return_data(arrs, [2, 1, 3])
        for j in range(3):
            ref = g["reply_%s_bytes_%d" % (tag, j)].tobytes()
            assert ours[j] == ref and orc[j] == ref


def test_oracle_encoders_reproduce_reference_bytes(g):
    # re-encoding what the reference decoded gives back the reference's bytes (forward states and training data)
    ids, msgs_bytes = [str(x) for x in g["fwd_ids"]], b""
    row = 0
    for pid_str, n in zip(ids, (2, 1, 3)):
        ip, pid = pid_str.split("_")
        msgs_bytes += R.easybytes_encode_forward_states(ip, int(pid), [g["fwd_slot_%d" % i][row:row + n] for i in range(4)])
        row += n
    assert msgs_bytes == g["fwd_bytes"].tobytes()
    st = [g["bwd_state_%d" % i] for i in range(2)]
    other = [g["bwd_other_%d" % i] for i in range(4)]
    assert R.easybytes_encode_backward_data(st, other, {"mean_reward": 1.5, "n": 3}) == g["bwd_bytes"].tobytes()


class _Logger:
    def __init__(self):
        self.rows, self.tags = [], None

    def update_tensor_tags(self, prefix, idx):
        self.tags = (prefix, idx)

    def add(self, value, key):
        self.rows.append((key, value))


def _configs(kind, B_env, redis_ns):
    spec = R.SPECS[kind]
    cfg = types.SimpleNamespace(
        PREDICTOR_REDIS_HOST=redis_ns, PREDICTOR_REDIS_PORT=1, MIDDLE_REDIS_HOST=redis_ns, MIDDLE_REDIS_PORT=2,
        TRAINER_REDIS_HOST=redis_ns, TRAINER_REDIS_PORT=3, TASK_NAME="t", PREDICTING_STATES_KEY="states",
        PRE_ACTIONS_KEY="act{}", UPDATE_TAG_KEY="upd", ENV_NUM_DICT_KEY="envs", TIME_OUT=1, PLAY_MODE=False,
        DEMONSTRATE_MODE=False, DEMONSTRATE_LOAD_PATH="", SYNC=False, TRAIN_LOCK_KEY="lock", TRAINING_DATA_KEY="train",
        LOG_LOSS_FREQUENCY=1, SAVE_MODELS=False, SAVE_FREQUENCY=100, SAVE_MODEL_PATH="", MIMIC_START=False,
        LOAD_CHECKPOINT=False, TEST=False, MODULE_KEY="MODEL", DEVICE="cuda")
    cnn = types.SimpleNamespace(MODULE_TENSOR_DTYPE=torch.float32, MODULE_NUMPY_DTYPE=np.float32, DEVICE="cuda",
                                ACTIONS_DIM=spec.act_dim, TRAINING_MIN_BATCH=2 * B_env, MODEL_TO_REDIS_FREQUENCY=2)
    return {"config": cfg, "config_nn": cnn, "config_env": {"batch_num_per_env": B_env, "agent_num_per_env": 1},
            "redis_factory": lambda h, p: _FakeRedis(h, p)}


def _make(kind, seed=11):
    from ddrl4nav_b200.runner import make_net
    spec = R.SPECS[kind]
    params = R.init_params(spec, seed=seed)
    net = make_net(kind, device=None)
    net.load_state_dict(params, strict=True)
    return net.to("cuda"), spec, params


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["pong", "navlaser"])
def test_forward_thread_one_tick(kind):
    from ddrl4nav_b200.server import ForwardThread
    B_env, n_env = 3, 2
    net, spec, params = _make(kind)
    cfgs = _configs(kind, B_env, "fwd-" + kind)
    mid = _FakeRedis("fwd-" + kind, 2)
    net.conn, net.model_key = mid, "tMODEL"                         # Basenn.conn / model_key (nn/base.py:28-33)
    net.nn2redis(mid.pipeline(), "tupd")                            # trainer published weights once
    flag = types.SimpleNamespace(value=b"0")
    th = ForwardThread(net, 0, _Logger(), None, flag, cfgs)
    # two env processes send float64 observations (what Pong's wrapper emits, warputils.py:300)
    states = [s.numpy() for s in R.synth_states(kind, B_env * n_env, seed=3)]
    wire = [s.astype(np.float64) if i == 0 else s for i, s in enumerate(states)]
    payload = b"".join(R.easybytes_encode_forward_states("10.0.0.7", pid, [w[j * B_env:(j + 1) * B_env] for w in wire])
                       for j, pid in enumerate((4, 9)))
    gauss = spec.dist == "gaussian"
    gen = torch.Generator().manual_seed(5)
    draw = torch.randn(B_env * n_env, spec.act_dim, generator=gen) if gauss else torch.rand(B_env * n_env, generator=gen)
    assert th.tick(payload, draw=draw.cuda()) == n_env
    pre = _FakeRedis("fwd-" + kind, 1)
    f32 = [torch.from_numpy(w.astype(np.float32)) for w in wire]
    wa, wlp, wv = R.forward_body(spec, params, f32, draw, play_mode=False)
    for j, pid in enumerate((4, 9)):
        item = pre.blpop("tact10.0.0.7_%d" % pid)
        assert item is not None
        a, lp, v = R.easybytes_decode_data(item[1])
        sl = slice(j * B_env, (j + 1) * B_env)
        assert a.dtype == np.float32 and v.shape == (1, B_env, 1)
        np.testing.assert_allclose(v[0, :, 0], wv.numpy().reshape(-1)[sl], rtol=1e-4, atol=1e-5)
        if gauss:
            np.testing.assert_allclose(a, wa.numpy()[sl], rtol=1e-4, atol=1e-5)
        else:
            assert (a == wa.numpy()[sl]).mean() >= 0.5                 # identical unless a draw sits within 1e-7 of a CDF edge
        np.testing.assert_allclose(lp[a == wa.numpy()[sl]] if not gauss else lp, (wlp.numpy()[sl])[a == wa.numpy()[sl]] if not gauss else wlp.numpy()[sl], rtol=1e-3, atol=1e-4)
    assert th.episode == 1 and th.logger_f.rows[0][0] == "ForwardTime-ms"


@pytest.mark.gpu
def test_backward_threads_one_batch():
    from ddrl4nav_b200.server import BackwardGetDataThread, BackwardTrainThread, BackwardQueue
    kind, B_env = "pong", 4
    net, spec, params = _make(kind)
    net.training_iter_time = 2
    net.model_key = "tMODEL"
    cfgs = _configs(kind, B_env, "bwd")
    flag = types.SimpleNamespace(value=b"0")
    q = BackwardQueue("cuda")
    getter = BackwardGetDataThread(net, q, 0, _Logger(), None, flag, cfgs)
    trainer = BackwardTrainThread(net, q, 0, _Logger(), None, flag, cfgs)
    train_redis, mid = _FakeRedis("bwd", 3), _FakeRedis("bwd", 2)
    states = R.synth_states(kind, 2 * B_env, seed=9)
    actions, old_logps, advs, returns = R.synth_learn_batch(spec, params, states, seed=2)
    for j in range(2):
        sl = slice(j * B_env, (j + 1) * B_env)
        other = [advs[sl].numpy(), actions[sl].numpy(), old_logps[sl].numpy(), returns[None, sl].numpy()]
        train_redis.lpush("ttrain", R.easybytes_encode_backward_data([states[0][sl].numpy().astype(np.float64)], other,
                                                                      {"reward": float(j)}))
    assert getter.get_train_data() and getter.get_train_data() and not getter.get_train_data()
    assert mid.get("tlock") == b"1"
    trainer.prepare()
    assert int(mid.get("tupd")) == 1 and mid.get("tMODEL") is not None
    iters = trainer.train_once()
    assert iters == 2 and mid.get("tlock") == b"0" and int(mid.get("tupd")) == 2       # update_time 2 -> nn2redis again
    keys = [k for k, _ in trainer.logger_f.rows]
    assert "reward" in keys and "PpoTotalLoss" in keys and trainer.data_len == 2 * B_env
    # the first logged loss = the oracle's first iteration on the same batch
    st = R.LearnState(spec, params)
    want, _, _ = R.learn_iteration(st, [states[0]], advs, actions, old_logps, returns, R.PPOHyper())
    got = {}
    for k, v in trainer.logger_f.rows:                        # first logged value of every key = iteration 1
        if k in ("ActorLoss", "VLoss", "EntLoss"):
            got.setdefault(k, v[0])
    for k in ("ActorLoss", "VLoss"):
        assert abs(got[k] - want[k]) <= 1e-4 * max(1.0, abs(want[k]))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["pong", "navlaser"])
def test_device_reply_encode_and_chunked_payload(kind):
    """step_bytes_replies (decode, engine, reply encode on the device) == host encoder over step_bytes' arrays, and a payload
    streamed in chunks of whole messages gives the same replies as the single-shot path."""
    from ddrl4nav_b200.server import ForwardModule, encode_forward_replies
    B_env, n_env = 5, 7
    net, spec, params = _make(kind)
    states = [s.numpy() for s in R.synth_states(kind, B_env * n_env, seed=4)]
    wire = [(s * 255).astype(np.uint8) if (i == 0 and kind == "pong") else s for i, s in enumerate(states)]
    payload = b"".join(R.easybytes_encode_forward_states("10.1.2.3", 100 + j, [w[j * B_env:(j + 1) * B_env] for w in wire])
                       for j in range(n_env))
    gen = torch.Generator().manual_seed(6)
    draw = (torch.randn(B_env * n_env, spec.act_dim, generator=gen) if spec.dist == "gaussian"
            else torch.rand(B_env * n_env, generator=gen)).cuda()
    fm = ForwardModule(net, device="cuda")
    ids, arrs = fm.step_bytes(payload, draw=draw)
    want = encode_forward_replies(arrs, [B_env] * n_env)
    ids2, got = fm.step_bytes_replies(payload, B_env, draw=draw)
    assert ids2 == ids and got == want
    # chunked: at most two messages per chunk; also from a pinned uint8 tensor (copied to the device from where it lies)
    fm_c = ForwardModule(net, device="cuda", chunk_bytes=2 * (len(payload) // n_env) + 8)
    ids3, got3 = fm_c.step_bytes_replies(payload, B_env, draw=draw)
    assert ids3 == ids
    pinned = torch.frombuffer(bytearray(payload), dtype=torch.uint8).pin_memory()
    ids4, got4 = fm_c.step_bytes_replies(pinned, B_env, draw=draw)
    assert ids4 == ids
    for a, b, c in zip(want, got3, got4):
        da, db, dc = (R.easybytes_decode_data(x) for x in (a, b, c))
        for u, v, w in zip(da, db, dc):
            assert u.shape == v.shape == w.shape
            np.testing.assert_allclose(v, u, rtol=1e-5, atol=1e-6)       # chunk amax differs: scaled-fp16 rounding, not bit-equal
            np.testing.assert_array_equal(v, w)
