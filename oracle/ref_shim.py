"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (DRL-Navigation/DDRL4NAV).

Only usable where ``/root/reference`` is mounted (the build container, never the GPU
box).  It is used by ``tests/golden/make_golden.py`` to generate the committed golden
vectors and by the ``not gpu`` tests to cross-check ``oracle/restate.py`` against the live
reference.  Nothing in ``ddrl4nav_b200/`` may import this file.

The reference needs a ``redis`` module at import time (``USTC_lab/nn/base.py:7``,
``USTC_lab/server/utils.py:9``); none is installed and there is no network, so a minimal
in-memory stand-in is injected into ``sys.modules`` (SURVEY.md section 8c lists exactly
which attributes the reference touches at import / constructor time).
"""
import os
import sys
import types
from types import SimpleNamespace

# Where the unmodified reference lives: the mounted tree in the build container, else the snapshot `build()` drops into
# the git-ignored oracle/_ref/ (python sources only; it travels to the GPU box like a built .so, oracle/snapshot_ref.py).
_HERE = os.path.dirname(os.path.abspath(__file__))
SNAPSHOT_ROOT = os.path.join(_HERE, "_ref")


def _default_root():
    env = os.environ.get("DDRL_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/USTC_lab"):
        return "/root/reference"
    return SNAPSHOT_ROOT


REFERENCE_ROOT = _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "USTC_lab"))


class _FakePipeline:
    def __init__(self, store):
        self._store = store
        self._ops = []

    def set(self, k, v):
        self._ops.append(("set", k, v)); return self

    def incr(self, k):
        self._ops.append(("incr", k)); return self

    def lpush(self, k, *v):
        self._ops.append(("lpush", k) + tuple(v)); return self

    def get(self, k):
        self._ops.append(("get", k)); return self

    def execute(self):
        out = []
        for op in self._ops:
            out.append(getattr(self._store, op[0])(*op[1:]))
        self._ops = []
        return out


class _FakeRedis:
    """dict-backed subset of redis.Redis (keys are shared per (host, port))."""
    _stores = {}

    def __init__(self, host=None, port=None, **kw):
        self._d = _FakeRedis._stores.setdefault((host, port), {})

    def pipeline(self):
        return _FakePipeline(self)

    def get(self, k):
        return self._d.get(k)

    def set(self, k, v):
        self._d[k] = v if isinstance(v, bytes) else str(v).encode(); return True

    def incr(self, k):
        v = int(self._d.get(k, b"0")) + 1
        self._d[k] = str(v).encode(); return v

    def hset(self, name, key=None, value=None, mapping=None):
        h = self._d.setdefault(name, {})
        if key is not None:
            h[key] = value
        if mapping:
            h.update(mapping)
        return 1

    def hgetall(self, name):
        return dict(self._d.get(name, {}))

    def lpush(self, k, *vals):
        lst = self._d.setdefault(k, [])
        for v in vals:
            lst.insert(0, v)
        return len(lst)

    def brpop(self, k, timeout=0):
        lst = self._d.get(k) or []
        return (k, lst.pop()) if lst else None

    def blpop(self, k, timeout=0):
        lst = self._d.get(k) or []
        return (k, lst.pop(0)) if lst else None


def install_fake_redis():
    if "redis" in sys.modules and not getattr(sys.modules["redis"], "_ddrl_fake", False):
        return  # a real redis package exists; leave it alone
    mod = types.ModuleType("redis")
    mod._ddrl_fake = True
    mod.Redis = _FakeRedis
    client = types.ModuleType("redis.client")
    client.Pipeline = _FakePipeline
    client.Redis = _FakeRedis
    mod.client = client
    sys.modules["redis"] = mod
    sys.modules["redis.client"] = client


def import_reference():
    """Returns the imported ``USTC_lab`` package of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    install_fake_redis()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import USTC_lab  # noqa: F401
    import USTC_lab.nn  # noqa: F401
    import USTC_lab.data  # noqa: F401
    import USTC_lab.server  # noqa: F401
    import USTC_lab.agent  # noqa: F401
    import USTC_lab.config  # noqa: F401
    return USTC_lab


def make_ref_configs(discrete: bool, n_act: int, share: bool = False, env_type="gym",
                     env_name="PongDeterministic-v4"):
    """BaseConfig / ConfigNN objects without the CLI (``runner/utils.py:50-56`` needs gym)."""
    import_reference()
    from USTC_lab.config import BaseConfig, ConfigNN
    if discrete:
        cnn = ConfigNN({"discrete_action": True, "discrete_actions": list(range(n_act))})
    else:
        cnn = ConfigNN({"discrete_action": False, "act_dim": n_act})
    cnn.SHARE_CNN_NET = share
    cnn.DEVICE = "cpu"
    parse = SimpleNamespace(th="h", tp=1, ph="h", pp=2, mh="h", mp=3, ch="h", cp=4,
                            model_dir="", ip="localhost", task="t")
    cfg = BaseConfig(parse, {"env_type": env_type, "env_name": env_name, "env_num": 4})
    cfg.DEVICE = "cpu"
    return cfg, cnn


def make_ref_net(kind: str):
    """Builds the reference PPO net the way ``runner/utils.py:59-170`` (create_net) does.

    kind: 'pong' (C1, unshared AtariPreNet x2, 6-way categorical),
          'navlaser' (C2, unshared NavPreNet1D x2, 2-d Gaussian),
          'navimg' (C5, shared NavPreNet, 28-way categorical),
          'navped' (shared NavPedPreNet on cat(map, 3-ch ped-map), 28-way categorical).
    """
    import_reference()
    from USTC_lab.nn import (PPO, AtariPreNet, NavPreNet, NavPedPreNet, NavPreNet1D, Critic)
    if kind == "pong":
        cfg, cnn = make_ref_configs(True, 6, share=False)
        pa, pc = AtariPreNet(4, last_output_dim=512, device="cpu"), AtariPreNet(4, last_output_dim=512, device="cpu")
        actor = cnn.ACTOR_CLASS(action_output_dim=6, device="cpu", soft_max_grid=True, last_input_dim=512,
                                pre=pa, nn_dtype=cnn.MODULE_TENSOR_DTYPE)
        critic = Critic(device="cpu", last_input_dim=512, pre=pc)
        return PPO(actor, critic, None, None, cfg, cnn), cfg, cnn
    if kind == "navlaser":
        cfg, cnn = make_ref_configs(False, 2, share=False, env_type="robot_nav")
        pa, pc = NavPreNet1D(image_channel=3, last_output_dim=512), NavPreNet1D(image_channel=3, last_output_dim=512)
        actor = cnn.ACTOR_CLASS(action_output_dim=2, device="cpu", soft_max_grid=True, last_input_dim=512,
                                nn_dtype=cnn.MODULE_TENSOR_DTYPE, pre=pa)
        critic = Critic(device="cpu", last_input_dim=512, pre=pc)
        return PPO(actor, critic, None, None, cfg, cnn), cfg, cnn
    if kind == "navimg":
        cfg, cnn = make_ref_configs(True, 28, share=True, env_type="robot_nav")
        actor = cnn.ACTOR_CLASS(action_output_dim=28, device="cpu", soft_max_grid=True, last_input_dim=512,
                                nn_dtype=cnn.MODULE_TENSOR_DTYPE)
        critic = Critic(device="cpu", last_input_dim=512)
        prenet = NavPreNet(image_channel=1, last_output_dim=512)
        return PPO(actor, critic, prenet, None, cfg, cnn), cfg, cnn
    if kind == "navped":
        cfg, cnn = make_ref_configs(True, 28, share=True, env_type="robot_nav")
        actor = cnn.ACTOR_CLASS(action_output_dim=28, device="cpu", soft_max_grid=True, last_input_dim=512,
                                nn_dtype=cnn.MODULE_TENSOR_DTYPE)
        critic = Critic(device="cpu", last_input_dim=512)
        prenet = NavPedPreNet(image_channel=4, last_output_dim=512)
        return PPO(actor, critic, prenet, None, cfg, cnn), cfg, cnn
    raise ValueError(kind)
