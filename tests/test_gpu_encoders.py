"""Stand-alone encoder / head calls of the reference's module API on the device:
``enc(states) -> [B, 512]`` (nn/atari_encoder.py:25-32, nn/nav_encoder.py:35-43,66-79,115-128), ``actor(states)`` with its
own encoder (nn/actor.py:26-40), ``critic(states)`` (nn/critic.py:14-21) -- against the oracle restatement."""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def _make(kind):
    from ddrl4nav_b200.runner import make_net
    spec = R.SPECS[kind]
    params = R.init_params(spec, seed=11)
    net = make_net(kind, device=None)
    net.load_state_dict(params, strict=True)
    return net.to("cuda"), spec, params


@pytest.mark.parametrize("kind,B", [("pong", 8), ("navimg", 6), ("navlaser", 4), ("navped", 5)])
def test_encoder_inside_ppo_returns_features(kind, B):
    net, spec, params = _make(kind)
    states = R.synth_states(kind, B, seed=5)
    dstates = [s.cuda() for s in states]
    with torch.no_grad():
        if spec.shared:
            want = R.encoder_forward(spec.arch, params, "prenet.", states)
            got = net.prenet(dstates)
            assert got.shape == (B, 512) and rel_err(got, want) < 1e-5
        else:
            for enc, pre in ((net.actor.pre, "actor.pre."), (net.critic.pre, "critic.pre.")):
                want = R.encoder_forward(spec.arch, params, pre, states)
                got = enc(dstates)
                assert got.shape == (B, 512) and rel_err(got, want) < 1e-5
            # the heads' own stand-alone calls chain through their encoders
            out = R.ppo_forward(spec, params, states)
            v = net.critic(dstates)
            assert v.shape == (B, 1) and rel_err(v.reshape(-1), out["values"].reshape(-1)) < 1e-5
            pi, _ = net.actor(dstates, play_mode=True)
            ref = out["probs"] if spec.dist == "categorical" else out["mu"]
            assert rel_err(pi, ref) < 1e-5


@pytest.mark.parametrize("kind,B", [("pong", 5), ("navimg", 3)])
def test_encoder_on_its_own(kind, B):
    from ddrl4nav_b200.nn import AtariPreNet, NavPreNet
    spec = R.SPECS[kind]
    params = R.init_params(spec, seed=3)
    pre = "prenet." if spec.shared else "actor.pre."
    enc = AtariPreNet(4, 512) if kind == "pong" else NavPreNet(image_channel=1)
    enc.load_state_dict({k[len(pre):]: v for k, v in params.items() if k.startswith(pre)}, strict=True)
    enc = enc.cuda()
    states = R.synth_states(kind, B, seed=8)
    with torch.no_grad():
        want = R.encoder_forward(spec.arch, params, pre, states)
        assert rel_err(enc([s.cuda() for s in states]), want) < 1e-5
        # weight edits are picked up by the next call
        enc.conv1.bias.add_(0.25)
        p2 = dict(params)
        p2[pre + "conv1.bias"] = params[pre + "conv1.bias"] + 0.25
        assert rel_err(enc([s.cuda() for s in states]), R.encoder_forward(spec.arch, p2, pre, states)) < 1e-5


def test_cpu_encoder_fails_loudly():
    from ddrl4nav_b200._lib import DDRLError
    from ddrl4nav_b200.nn import AtariPreNet
    with pytest.raises(DDRLError):
        AtariPreNet(4, 512)([torch.zeros(1, 4, 84, 84)])
