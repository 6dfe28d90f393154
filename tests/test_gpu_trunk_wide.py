"""GPU tests of the wide fused tcgen05 trunk (csrc/trunk_wide.cu; 32 / 64 filters, AlphaSame and
BaseResNet / AuxBaseResNet).  Floating point: compared with a plain PyTorch fp32 reference of the same
op (the reference's own modules, architectures.py:27-353), tolerance written in each test."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _randomise_bn(net):
    import torch
    for m in net.modules():  # non-trivial BatchNorm statistics
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
            m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.7, 1.3); m.bias.data.normal_(0, 0.2)
    return net


def _net(family, blocks, filters, seed):
    import torch
    from tetris_reinforcement_learning_b200 import architectures as arch
    torch.manual_seed(seed)
    if family == "alphasame":
        net = arch.AlphaSame(arch.AlphaSameConfig(blocks=blocks, filters=filters))
    elif family == "base":
        net = arch.BaseResNet(arch.BaseResNetConfig(blocks=blocks, filters=filters))
    else:
        net = arch.AuxBaseResNet(arch.AuxBaseResNetConfig(blocks=blocks, filters=filters))
    return _randomise_bn(net.to("cuda:0").eval())


def _reference_trunk(net, grids):
    """fp32 PyTorch: what one row of the kernel's output holds."""
    import torch
    from tetris_reinforcement_learning_b200 import architectures as arch
    from tetris_reinforcement_learning_b200.trunk import _fold_bn
    with torch.no_grad():
        if isinstance(net, arch.AlphaSame):
            return net.grid_features(grids)
        feat = net._process_grid(grids)
        so, _ = _fold_bn(net.own_collapse[1])
        own = net.own_collapse[0](feat) * so[None, :, None, None]          # bn scale applied, bias joins the FiLM term
        opp = net.opp_collapse(feat)
        return torch.cat([own, opp], dim=1).flatten(1)


CASES = [("alphasame", 1, 32, 3), ("alphasame", 1, 64, 1), ("alphasame", 2, 64, 700), ("alphasame", 20, 64, 900),
         ("alphasame", 10, 32, 1000), ("base", 1, 32, 7), ("base", 8, 32, 2000), ("aux", 8, 32, 889), ("base", 3, 64, 500)]


@pytest.mark.parametrize("family,blocks,filters,n", CASES)
def test_wide_trunk_matches_pytorch_fp32(family, blocks, filters, n):
    """Tolerance: bf16 operands AND a bf16 residual stream with fp32 accumulation ->
    |err| <= 0.04 * max|ref| + 0.03 elementwise and mean |err| <= 1.5 % of mean |ref|."""
    import torch
    from tetris_reinforcement_learning_b200 import trunk_wide
    net = _net(family, blocks, filters, 1)
    assert trunk_wide.supports(net)
    g = torch.Generator(device="cuda:0").manual_seed(2)
    grids = (torch.rand((n, 1, 40, 10), generator=g, device="cuda:0") < 0.35).float()
    grids[0] = 0          # empty board
    grids[-1] = 1         # full board
    ref = _reference_trunk(net, grids)
    wt = trunk_wide.WideTrunk(trunk_wide.pack_wide_trunk(net), "cuda:0")
    got = trunk_wide.wide_trunk_forward(wt, grids.to(torch.bfloat16)).float()
    torch.cuda.synchronize()
    wt.check()
    assert got.shape == ref.shape
    err = (got - ref).abs()
    scale = ref.abs().max().item()
    assert torch.isfinite(got).all()
    assert err.max().item() <= 0.04 * scale + 0.03, (err.max().item(), scale)
    assert err.mean().item() <= 0.015 * ref.abs().mean().item() + 1e-3, (err.mean().item(), ref.abs().mean().item())
    # a second launch re-uses the scratch and the work counters: same answer
    got2 = trunk_wide.wide_trunk_forward(wt, grids.to(torch.bfloat16)).float()
    torch.cuda.synchronize()
    assert torch.equal(got, got2)


@pytest.mark.parametrize("family,blocks,filters", [("alphasame", 4, 64), ("aux", 8, 32), ("base", 2, 64)])
def test_wide_evaluator_matches_module(family, blocks, filters):
    """Whole net: value within 0.03, logits within 5 % of max |logit| + 0.05 of the fp32 module."""
    import torch
    from tetris_reinforcement_learning_b200 import trunk_wide
    net = _net(family, blocks, filters, 3)
    B = 200
    g = torch.Generator(device="cuda:0").manual_seed(5)
    grids = (torch.rand((2 * B, 1, 40, 10), generator=g, device="cuda:0") < 0.3).float()
    extras = torch.randint(0, 2, (B, 105), generator=g, device="cuda:0").float()
    with torch.no_grad():
        out = net.forward_packed(grids, extras)
    v_ref, l_ref = out[0], out[1]
    ev = trunk_wide.make_wide_evaluator(copy.deepcopy(net))
    with torch.no_grad():
        v, l = ev(grids.to(torch.bfloat16), extras.to(torch.bfloat16))
    torch.cuda.synchronize()
    ev.trunk.check()
    assert (v.float().reshape(-1) - v_ref.reshape(-1)).abs().max().item() < 0.03
    assert l.shape[1] >= 11583
    assert (l[:, :11583].float() - l_ref).abs().max().item() < 0.05 * l_ref.abs().max().item() + 0.05


@pytest.mark.parametrize("family,blocks,filters", [("aux", 2, 32), ("alphasame", 2, 64)])
def test_wide_engine_feature_reuse_is_exact(family, blocks, filters):
    """Self-play with the wide evaluator inside the engine's CUDA graph; the engine with trunk-feature reuse
    must produce bit-identical searches to the engine that sends both boards of every leaf through the trunk."""
    import torch
    from tetris_reinforcement_learning_b200 import architectures as arch, trunk_wide
    from tetris_reinforcement_learning_b200.config import Config
    from tetris_reinforcement_learning_b200.selfplay import SelfPlayEngine, best_evaluator
    net = _net(family, blocks, filters, 7)
    ev = best_evaluator(copy.deepcopy(net))
    assert getattr(ev, "cached", None) is not None and isinstance(ev.cached, trunk_wide.CachedWideEvaluator)
    cfg = Config(visual=False, ruleset="s2", model="pytorch", model_config=arch.AlphaSameConfig(), MAX_ITER=12, training=True)
    out = []
    for reuse in (True, False):
        eng = SelfPlayEngine(cfg, ev, 96, seed=3, feature_dtype=torch.bfloat16, max_rounds=4, reuse_trunk_features=reuse,
                             gather_policy=False)   # same dense bf16 policy head on both sides: the comparison is about the trunk
        assert (eng.cached_eval is not None) == reuse
        eng.step(300)
        samples, ends = eng.drain()
        ev.trunk.check()
        out.append((samples, ends, eng.get_ctl()))
    (s0, e0, c0), (s1, e1, c1) = out
    assert len(s0) > 50 and len(s0) == len(s1) and len(e0) == len(e1) and len(e0) > 0
    s0, s1 = (np.sort(s, order=["game_id", "search_no"]) for s in (s0, s1))
    for name in s0.dtype.names:
        assert np.array_equal(s0[name], s1[name]), name
    assert (c0["status"] == 0).all() and (c1["status"] == 0).all()
