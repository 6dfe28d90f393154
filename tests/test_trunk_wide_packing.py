"""CPU check of the host side of the wide fused trunk (trunk_wide.pack_wide_trunk): a plain fp32 emulation of
what csrc/trunk_wide.cu computes FROM THE PACKED TENSORS ONLY (row-Toeplitz B matrices in UMMA core-matrix
order, stem tables, folded BatchNorm slots, head rows) must reproduce the PyTorch modules
(reference architectures.py:27-353).  The kernel itself is compared with PyTorch in test_gpu_trunk_wide.py."""
import pytest
import torch

from tetris_reinforcement_learning_b200 import architectures as arch, trunk_wide
from tetris_reinforcement_learning_b200.trunk import _fold_bn


def _randomise_bn(net):
    for m in net.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
            m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.7, 1.3); m.bias.data.normal_(0, 0.2)
    return net


def _emulate(p, grids):
    """grids [n,40,10] 0/1 -> [n, n_out*400], following the kernel's dataflow in fp32."""
    f, taps, post, n_out = p["filters"], p["stem_taps"], p["post_act"], p["n_out"]
    L = 2 * p["n_blocks"]
    n = grids.shape[0]
    consts = p["consts"].float()
    slots = consts[:(L + 1) * 3 * f].reshape(L + 1, 3, f)
    head = consts[(L + 1) * 3 * f:]
    hw, hs, hb = head[:n_out * f].reshape(n_out, f), head[n_out * f:n_out * f + n_out], head[n_out * f + n_out:]
    lut = p["stem_lut"].float()                                  # [dy][pat][f]
    half = taps // 2
    padded = torch.zeros((n, 40 + 2 * half, 10 + 2 * half))
    padded[:, half:half + 40, half:half + 10] = grids
    acc = torch.zeros((n, 40, 10, f))
    for dy in range(taps):
        pat = torch.zeros((n, 40, 10), dtype=torch.long)
        for i in range(taps):
            pat += (padded[:, dy:dy + 40, i:i + 10] > 0).long() << i
        acc += lut[dy][pat]
    if post:
        x = torch.relu(acc + slots[0, 0])
        t = x
    else:
        x = acc
        t = torch.relu(slots[0, 1] * x + slots[0, 2])
    # B matrices back from the core-matrix order: [layer][dy][kc][ng][n][k] -> [layer][dy][(j, oc)][ic]
    w = p["w_packed"].float().reshape(L, 3, f // 8, 3 * f // 8, 8, 8).permute(0, 1, 3, 4, 2, 5).reshape(L, 3, 3, f, f)

    def conv(a, wl):   # a [n,40,10,f] ; D[r, x_out] += A[r + dy - 1, x_in] B_dy[(x_out - x_in + 1, oc), ic]
        ap = torch.zeros((n, 42, 10, f))
        ap[:, 1:41] = a
        out = torch.zeros((n, 40, 10, f))
        for dy in range(3):
            rows = ap[:, dy:dy + 40]
            for j in range(3):
                contrib = torch.einsum("nrxi,oi->nrxo", rows, wl[dy, j])      # input column x -> output column x - 1 + j
                lo, hi = max(0, 1 - j), min(10, 11 - j)                        # input columns whose target is on the board
                out[:, :, lo - 1 + j:hi - 1 + j] += contrib[:, :, lo:hi]
        return out

    for l in range(L):
        a = conv(t, w[l]) + slots[l + 1, 0]
        if l % 2 == 0:
            t = torch.relu(a)
        elif post:
            x = torch.relu(x + a)
            t = x
        else:
            x = x + a
            t = torch.relu(slots[l + 1, 1] * x + slots[l + 1, 2])
    o = torch.einsum("nrxc,kc->nkrx", t, hw) * hs[None, :, None, None] + hb[None, :, None, None]
    if post:
        o = torch.cat([o[:, :4], torch.relu(o[:, 4:])], dim=1)
    else:
        o = torch.relu(o)
    return o.reshape(n, n_out * 400)


@pytest.mark.parametrize("family,blocks,filters", [("alphasame", 2, 32), ("alphasame", 1, 64), ("base", 2, 32), ("aux", 1, 64)])
def test_packed_wide_trunk_reproduces_the_module(family, blocks, filters):
    torch.manual_seed(11)
    if family == "alphasame":
        net = arch.AlphaSame(arch.AlphaSameConfig(blocks=blocks, filters=filters))
    elif family == "base":
        net = arch.BaseResNet(arch.BaseResNetConfig(blocks=blocks, filters=filters))
    else:
        net = arch.AuxBaseResNet(arch.AuxBaseResNetConfig(blocks=blocks, filters=filters))
    net = _randomise_bn(net.eval())
    assert trunk_wide.supports(net)
    p = trunk_wide.pack_wide_trunk(net, device="cpu")
    grids = (torch.rand((5, 1, 40, 10)) < 0.35).float()
    grids[0] = 0
    grids[-1] = 1
    with torch.no_grad():
        if family == "alphasame":
            ref = net.grid_features(grids)
        else:
            feat = net._process_grid(grids)
            so, _ = _fold_bn(net.own_collapse[1])
            ref = torch.cat([net.own_collapse[0](feat) * so[None, :, None, None], net.opp_collapse(feat)], dim=1).flatten(1)
    got = _emulate(p, grids[:, 0])
    # packed weights are bf16: 3 significant digits per weight, errors average out over K = 9 F
    assert (got - ref).abs().max().item() <= 0.03 * ref.abs().max().item() + 0.02
    assert (got - ref).abs().mean().item() <= 0.01 * ref.abs().mean().item() + 1e-3


def test_supports_rejects_other_shapes():
    assert not trunk_wide.supports(arch.AlphaSame(arch.AlphaSameConfig(blocks=1, filters=16)))
    assert not trunk_wide.supports(arch.AlphaSame(arch.AlphaSameConfig(blocks=1, filters=32, kernels=2)))
    assert not trunk_wide.supports(arch.AlphaSame(arch.AlphaSameConfig(blocks=1, filters=32, value_head_neurons=32)))
    assert not trunk_wide.supports(arch.BaseResNet(arch.BaseResNetConfig(blocks=1, filters=48)))
    assert not trunk_wide.supports(arch.BaseResNet(arch.BaseResNetConfig(blocks=1, filters=32, own_kernels=2)))
    assert trunk_wide.supports(arch.AuxBaseResNet(arch.AuxBaseResNetConfig()))


@pytest.mark.parametrize("family", ["base", "aux"])
def test_baseresnet_heads_reproduce_the_module(family):
    """BaseResNetHeads (FiLM-add pushed through the linear own collapse, BatchNorm folded) on exact fp32 trunk
    outputs == BaseResNet.forward_packed (reference architectures.py:235-271)."""
    torch.manual_seed(5)
    cfg = arch.BaseResNetConfig(blocks=1, filters=32) if family == "base" else arch.AuxBaseResNetConfig(blocks=1, filters=32)
    net = _randomise_bn((arch.BaseResNet if family == "base" else arch.AuxBaseResNet)(cfg).eval())
    b = 6
    grids = (torch.rand((2 * b, 1, 40, 10)) < 0.35).float()
    extras = torch.randint(0, 3, (b, 105)).float()
    with torch.no_grad():
        out = net.forward_packed(grids, extras)
        feat = net._process_grid(grids)
        so, _ = _fold_bn(net.own_collapse[1])
        rows = torch.cat([net.own_collapse[0](feat) * so[None, :, None, None], net.opp_collapse(feat)], dim=1).flatten(1)
        heads = trunk_wide.BaseResNetHeads(net, dtype=torch.float32)
        v, l = heads(rows[:b], rows[b:], extras)
    assert torch.allclose(v.reshape(-1), out[0].reshape(-1), atol=1e-5)
    assert torch.allclose(l[:, :11583], out[1], atol=1e-4, rtol=1e-4)
