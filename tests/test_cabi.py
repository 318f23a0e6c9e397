"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/ddrl_b200.h declares; the engine's parameter table equals the reference's
named_parameters() order (oracle.param_table, itself pinned to the live reference); the Python
mirror registers parameters in the same order; the weight wire format round-trips."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest
import torch

from oracle import ref_shim
from oracle import restate as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from ddrl4nav_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    return _lib


def test_library_exports_every_declared_symbol():
    _lib = _ensure_built()
    hdr = open(os.path.join(ROOT, "include", "ddrl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ddrl_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "missing export: " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    assert lib.ddrl_version() >= 100
    assert lib.ddrl_error_string(-1) == b"bad argument"


@pytest.mark.parametrize("kind", ["pong", "navlaser", "navimg", "navped"])
def test_engine_param_table_is_reference_order(kind):
    _lib = _ensure_built()
    lib = _lib.load()
    spec = R.SPECS[kind]
    desc = _lib.NetDesc(_lib.ARCH[spec.arch], spec.in_ch, spec.act_dim, _lib.DIST[spec.dist], int(spec.shared), spec.feat, 0, 0)
    h = C.c_void_p()
    assert lib.ddrl_net_create(C.byref(desc), C.byref(h)) == 0
    table = R.param_table(spec)
    assert lib.ddrl_net_num_tensors(h) == len(table)
    name = C.create_string_buffer(128)
    shape = (C.c_int64 * 4)()
    ndim, off = C.c_int(), C.c_int64()
    total = 0
    for i, (n, shp) in enumerate(table):
        assert lib.ddrl_net_tensor_info(h, i, name, 128, shape, C.byref(ndim), C.byref(off)) == 0
        assert name.value.decode() == n
        assert tuple(shape[k] for k in range(ndim.value)) == tuple(shp)
        assert off.value == total
        total += int(np.prod(shp))
    assert lib.ddrl_net_num_params(h) == total == {"pong": 3371847, "navlaser": 12799685, "navimg": 5633565, "navped": 5635293}[kind]
    # argument checking without a GPU
    assert lib.ddrl_net_backward(h, None, 0, 1, 1, None, None, None, None, None, 0, None) != 0
    assert lib.ddrl_net_destroy(h) == 0
    bad = _lib.NetDesc(99, 1, 1, 0, 0, 512, 0, 0)
    assert lib.ddrl_net_create(C.byref(bad), C.byref(h)) == -1


def test_laser_channel_variant_tables_agree():
    """The NON-reference 3 x 960 laser variant (SURVEY 8d asks for it beside C2): the engine (ddrl_net_desc.laser_ch), the Python
    mirror and the oracle agree on the parameter table; laser_ch = 0 / 1 is the reference's Conv1d(1, 32, 5, 2)."""
    _lib = _ensure_built()
    lib = _lib.load()
    from ddrl4nav_b200.runner import make_net
    spec = R.SPECS["navlaser3"]
    assert spec.laser_ch == 3 and R.SPECS["navlaser"].laser_ch == 1
    table = R.param_table(spec)
    assert ("actor.pre.conv1d1.weight", (32, 3, 5)) in table
    net = make_net("navlaser3", device=None)
    assert [(n, tuple(p.shape)) for n, p in net.named_parameters()] == [(n, tuple(s)) for n, s in table]
    name = C.create_string_buffer(128)
    shape = (C.c_int64 * 4)()
    ndim, off = C.c_int(), C.c_int64()
    for laser_ch, want in ((3, table), (0, R.param_table(R.SPECS["navlaser"])), (1, R.param_table(R.SPECS["navlaser"]))):
        desc = _lib.NetDesc(_lib.ARCH[spec.arch], spec.in_ch, spec.act_dim, _lib.DIST[spec.dist], 0, spec.feat, 0, laser_ch)
        h = C.c_void_p()
        assert lib.ddrl_net_create(C.byref(desc), C.byref(h)) == 0
        assert lib.ddrl_net_num_tensors(h) == len(want)
        for i, (n, shp) in enumerate(want):
            assert lib.ddrl_net_tensor_info(h, i, name, 128, shape, C.byref(ndim), C.byref(off)) == 0
            assert name.value.decode() == n and tuple(shape[k] for k in range(ndim.value)) == tuple(shp)
        assert lib.ddrl_net_obs_elems(h, 0) == 960 * max(laser_ch, 1)
        assert lib.ddrl_net_destroy(h) == 0


def test_bind_host_to_device_is_harmless_without_a_gpu():
    import os
    from ddrl4nav_b200.dist import bind_host_to_device
    before = os.sched_getaffinity(0)
    assert bind_host_to_device(0) is None or os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)


@pytest.mark.parametrize("kind", ["pong", "navlaser", "navimg", "navped"])
def test_mirror_modules_register_reference_order(kind):
    from ddrl4nav_b200.runner import make_net
    net = make_net(kind, device=None)
    got = [(n, tuple(p.shape)) for n, p in net.named_parameters()]
    assert got == [(n, tuple(s)) for n, s in R.param_table(R.SPECS[kind])]
    if kind == "navlaser":
        assert float(net.actor.log_std[0]) == -0.5


def test_wire_format_roundtrip_cpu():
    from ddrl4nav_b200.runner import make_net
    net = make_net("mlp", device=None)
    blob = net.model_bytes()
    expect = b"".join(struct.pack(">I", p.dim()) + struct.pack(">%dI" % p.dim(), *p.shape) + p.detach().numpy().tobytes()
                      for _, p in net.named_parameters())
    assert blob == expect
    net2 = make_net("mlp", device=None)
    net2.load_model_bytes(blob)
    for (_, a), (_, b) in zip(net.named_parameters(), net2.named_parameters()):
        assert torch.equal(a, b)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not mounted")
def test_wire_format_matches_live_reference():
    """Our blob is byte-identical to the reference's nn2redis payload and loads into the reference net."""
    from ddrl4nav_b200.runner import make_net
    ref, _, _ = ref_shim.make_ref_net("navimg")
    params = R.init_params(R.SPECS["navimg"], seed=5)
    ref.load_state_dict(params)
    ours = make_net("navimg", device=None)
    ours.load_state_dict(params)

    class Pipe:
        def __init__(self): self.kv = {}
        def set(self, k, v): self.kv[k] = v
        def incr(self, k): self.kv[k] = self.kv.get(k, 0) + 1
        def execute(self): pass
        def get(self, k): return self.kv[k]
    p1, p2 = Pipe(), Pipe()
    ref.nn2redis(p1, "tag")
    ours.nn2redis(p2, "tag")
    assert p1.kv[ref.model_key] == p2.kv[ours.model_key]
    assert p2.kv["tag"] == 1
    ref.updatenn_by_redis(p2, ours.model_key)      # reference decodes our blob


def test_no_gpu_means_loud_failure():
    from ddrl4nav_b200 import DDRLError, kernels
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(DDRLError):
        kernels.gae(torch.zeros(3, 1, 4), torch.zeros(2, 1, 4), torch.zeros(2, 1, 4, dtype=torch.uint8), [0.99], 0.95)


def test_actors_are_instances_of_the_reference_classes_when_hosted_by_it():
    """server/forward.py:140-142 dispatches on isinstance(net.actor, GaussionActor / CategoricalActor) with the REFERENCE's
    classes: when the reference is the host application (imported first) our actors derive from them -- no patching."""
    import subprocess
    import sys
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference not mounted")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle import ref_shim; ref_shim.import_reference()\n"
        "import USTC_lab.nn as R\n"
        "from ddrl4nav_b200.nn import GaussionActor, CategoricalActor\n"
        "g = GaussionActor(action_output_dim=2, last_input_dim=512); c = CategoricalActor(6, last_input_dim=512)\n"
        "assert isinstance(g, R.GaussionActor) and isinstance(c, R.CategoricalActor)\n"
        "assert not isinstance(g, R.CategoricalActor) and not isinstance(c, R.GaussionActor)\n"
        "assert [n for n, _ in g.named_parameters()] == ['log_std', 'actor_linear.weight', 'actor_linear.bias']\n"
        "assert type(g).forward is not R.GaussionActor.forward and type(g).__mro__[1].__module__.startswith('ddrl4nav_b200')\n"
        "print('ok')\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
