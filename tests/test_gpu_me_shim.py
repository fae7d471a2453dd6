"""The MinkowskiEngine / torchac drop-in modules (linr_pcgc_b200.shim) against the CPU oracle: the L2->L1 boundary of
SURVEY.md 8(b).  Each test mirrors a call pattern of the reference's own modules (file:line in the docstrings)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-4, atol=1e-5)   # fp32 tolerance of the north star (1e-4 relative)


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import linr_pcgc_b200.shim as shim
    from linr_pcgc_b200 import _lib, synth
    from oracle import linr_oracle as O
    _lib.load()
    shim.install(force=True)
    import MinkowskiEngine as ME
    import torchac
    pts = synth.make_sequence("tiny", 1)[0].numpy()
    fr = O.prepare_frame(pts, None, 64)
    coord = fr["scales"][0]["coord"].astype(np.int32)            # sorted unique parents of scale 0
    nbr = torch.from_numpy(O.nbr27(coord).astype(np.int64))
    return ME, torchac, O, coord, nbr


def _sparse(ME, coord, feats):
    """generate_sparse (models/function_utils.py:13-18)."""
    xyz = torch.from_numpy(coord).cuda()
    C, F = ME.utils.sparse_collate([xyz.int()], [feats.float()])
    st = ME.SparseTensor(features=F, coordinates=C, tensor_stride=1, device=xyz.device)
    assert int((st.C[:, 1:] != xyz).sum()) == 0                   # function_utils.py:17
    return st, xyz


@pytest.mark.parametrize("cin,cout", [(1, 8), (3, 8), (4, 4), (7, 8), (8, 4), (8, 8)])
def test_conv3_forward_backward_match_oracle(env, cin, cout):
    """ME.MinkowskiConvolution(kernel_size=3) (models/upsample.py:17,90,95; models/resnet.py:15-51), every channel
    pair the live network uses (ConvA of the LDFE blocks has Cin = 1..7, models/upsample.py:72-76)."""
    ME, _, O, coord, nbr = env
    g = torch.Generator().manual_seed(100 * cin + cout)
    x = torch.randn(len(coord), cin, generator=g)
    conv = ME.MinkowskiConvolution(cin, cout, kernel_size=3, stride=1, bias=True, dimension=3).cuda()
    W, b = conv.kernel.detach().cpu().clone().requires_grad_(True), conv.bias.detach().cpu().clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = O.conv27(xr, nbr, W, b)
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    xs = x.cuda().requires_grad_(True)
    st, _ = _sparse(ME, coord, xs)
    out = conv(st)
    assert out.coordinate_map_key == st.coordinate_map_key and out.F.shape == (len(coord), cout)
    out.F.backward(dy.cuda())
    np.testing.assert_allclose(out.F.detach().cpu().numpy(), ref.detach().numpy(), **TOL)
    np.testing.assert_allclose(xs.grad.cpu().numpy(), xr.grad.numpy(), **TOL)
    np.testing.assert_allclose(conv.kernel.grad.cpu().numpy(), W.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(conv.bias.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-4, atol=1e-4)
    # no floating-point atomics: a second backward gives the same bits
    conv.zero_grad()
    xs2 = x.cuda().requires_grad_(True)
    st2, _ = _sparse(ME, coord, xs2)
    conv(st2).F.backward(dy.cuda())
    assert torch.equal(xs2.grad, xs.grad)


def test_conv_with_explicit_coordinates_and_kernel1(env):
    """ConvWithPrune calls conv(x, coords) with the input's own set (models/upsample.py:20-23,191); kernel_size 1 is a
    2-D [Cin,Cout] kernel (checkpoint contract)."""
    ME, _, O, coord, nbr = env
    x = torch.randn(len(coord), 8)
    st, _ = _sparse(ME, coord, x.cuda())
    conv = ME.MinkowskiConvolution(8, 8, kernel_size=3, stride=1, bias=True, dimension=3).cuda()
    a, b = conv(st), conv(st, st.C)
    assert torch.equal(a.F, b.F) and torch.equal(b.C, st.C)               # rows stay in the given order
    c = conv(st, st.C.clone())                                            # same content under another tensor
    assert torch.equal(c.F, a.F)
    with pytest.raises(NotImplementedError):
        conv(st, st.C[: len(coord) // 2])
    pw = ME.MinkowskiConvolution(8, 4, kernel_size=1, stride=1, bias=True, dimension=3).cuda()
    assert tuple(pw.kernel.shape) == (8, 4) and tuple(pw.bias.shape) == (1, 4)
    ref = O.conv1(x, pw.kernel.detach().cpu(), pw.bias.detach().cpu())
    np.testing.assert_allclose(pw(st).F.detach().cpu().numpy(), ref.numpy(), **TOL)
    for bad in (dict(stride=2), dict(dilation=2), dict(kernel_size=5), dict(dimension=2)):
        kw = dict(kernel_size=3, stride=1, dilation=1, bias=True, dimension=3)
        kw.update(bad)
        with pytest.raises(NotImplementedError):
            ME.MinkowskiConvolution(8, 8, **kw)


def test_block_composition_matches_oracle(env):
    """CNP.make_block + InceptionResNet.forward (models/upsample.py:88-97, models/resnet.py:55-60) assembled from
    the shim's modules, parameters named as in the reference's state_dict."""
    ME, _, O, coord, nbr = env
    S = 3
    sd = O.init_params(S, seed=21)
    pre = "upsampler.block_in"
    irn = f"{pre}.2.layers.0"

    def mk(name, cin, cout, k):
        m = ME.MinkowskiConvolution(cin, cout, kernel_size=k, stride=1, bias=True, dimension=3).cuda()
        with torch.no_grad():
            assert m.kernel.shape == sd[f"{name}.kernel"].shape and m.bias.shape == sd[f"{name}.bias"].shape
            m.kernel.copy_(sd[f"{name}.kernel"])
            m.bias.copy_(sd[f"{name}.bias"])
        return m

    A, B = mk(f"{pre}.0", 8, 8, 3), mk(f"{pre}.3", 8, 8, 3)
    c00, c01 = mk(f"{irn}.conv0_0", 8, 4, 3), mk(f"{irn}.conv0_1", 4, 4, 3)
    c10, c11, c12 = mk(f"{irn}.conv1_0", 8, 4, 1), mk(f"{irn}.conv1_1", 4, 4, 3), mk(f"{irn}.conv1_2", 4, 4, 1)
    relu = ME.MinkowskiReLU(inplace=True)
    x = torch.randn(len(coord), 8, generator=torch.Generator().manual_seed(5))
    st, _ = _sparse(ME, coord, x.cuda())
    y = relu(A(st))
    out0 = c01(relu(c00(y)))
    out1 = c12(relu(c11(relu(c10(y)))))
    z = ME.cat(out0, out1) + y                                            # models/resnet.py:58
    got = B(z)
    ref = O.block_forward(sd, pre, x, nbr)
    np.testing.assert_allclose(got.F.detach().cpu().numpy(), ref.detach().numpy(), **TOL)


def test_merge_prune_cat_union(env):
    """merge_two_frames = zero-pad + sparse add on one manager (models/function_utils.py:58-69); MinkowskiPruning
    (models/upsample.py:116); union of different sets."""
    ME, _, O, coord, nbr = env
    n = len(coord)
    f1, f2 = torch.randn(n, 8).cuda(), torch.randn(n, 3).cuda()
    s1, xyz = _sparse(ME, coord, f1)
    s2, _ = _sparse(ME, coord, f2)
    z1 = ME.SparseTensor(torch.cat([s1.F, torch.zeros((n, 3), device="cuda")], dim=-1), coordinates=s1.C, device=s1.device)
    z2 = ME.SparseTensor(torch.cat([torch.zeros((n, 8), device="cuda"), s2.F], dim=-1), coordinates=s2.C,
                         coordinate_manager=z1.coordinate_manager, device=s1.device)
    merged = z1 + z2
    assert torch.equal(merged.C, s1.C) and torch.equal(merged.F, torch.cat([f1, f2], dim=1))
    # pruning keeps row order; an all-true mask is an identity copy (instage = 1, models/upsample.py:120-124)
    prune = ME.MinkowskiPruning()
    full = prune(s1, torch.ones(n, dtype=torch.bool, device="cuda"))
    assert torch.equal(full.F, s1.F) and torch.equal(full.C, s1.C)
    mask = torch.zeros(n, dtype=torch.bool, device="cuda")
    mask[::3] = True
    sub = prune(s1, mask)
    assert torch.equal(sub.C[:, 1:], xyz[mask]) and torch.equal(sub.F, f1[mask])
    # a conv on the pruned set uses the pruned set's own neighbourhoods
    conv = ME.MinkowskiConvolution(8, 8, kernel_size=3, stride=1, bias=False, dimension=3).cuda()
    sub_nbr = torch.from_numpy(O.nbr27(coord[mask.cpu().numpy()]).astype(np.int64))
    ref = O.conv27(f1[mask].cpu(), sub_nbr, conv.kernel.detach().cpu(), None)
    np.testing.assert_allclose(conv(sub).F.detach().cpu().numpy(), ref.numpy(), **TOL)
    # union of two different sets on one manager: x-major sorted whatever the operand order, so it can be convolved
    a = ME.SparseTensor(f1[: n // 2], coordinates=s1.C[: n // 2], device="cuda")
    b = ME.SparseTensor(f1[n // 4:], coordinates=s1.C[n // 4:], coordinate_manager=a.coordinate_manager, device="cuda")
    want = torch.zeros_like(f1)
    want[: n // 2] += f1[: n // 2]
    want[n // 4:] += f1[n // 4:]
    for u in (a + b, b + a):
        assert torch.equal(u.C, s1.C)
        assert torch.allclose(u.F, want)
    ref_u = O.conv27(want.cpu(), nbr, conv.kernel.detach().cpu(), None)
    np.testing.assert_allclose(conv(b + a).F.detach().cpu().numpy(), ref_u.numpy(), **TOL)
    with pytest.raises(ValueError):
        s1 + s2                                                           # different managers
    with pytest.raises(ValueError):
        ME.cat(s1, s2)


def test_coordinate_sets_are_validated(env):
    """The compact kernel map needs ONE batch of x-major sorted unique rows: anything else raises instead of gathering
    wrong rows (the check runs once per coordinate set, when its tables are built)."""
    ME, _, O, coord, nbr = env
    conv = ME.MinkowskiConvolution(8, 8, kernel_size=3, stride=1, bias=False, dimension=3).cuda()
    n = len(coord)
    x = torch.randn(n, 8).cuda()
    perm = torch.randperm(n)
    C0 = torch.cat([torch.zeros((n, 1), dtype=torch.int32), torch.from_numpy(coord)], dim=1).cuda()
    with pytest.raises(ValueError, match="sorted"):
        conv(ME.SparseTensor(x, coordinates=C0[perm.cuda()].contiguous()))
    dup = torch.cat([C0[:5], C0[4:]], dim=0)
    with pytest.raises(ValueError, match="duplicate"):
        conv(ME.SparseTensor(torch.randn(n + 1, 8).cuda(), coordinates=dup.contiguous()))
    two = C0.clone()
    two[n // 2:, 0] = 1
    with pytest.raises(NotImplementedError, match="batch"):
        conv(ME.SparseTensor(x, coordinates=two))
    conv(ME.SparseTensor(x, coordinates=C0))       # the sorted unique single-batch set is accepted


def test_kernel_map_cache_follows_the_coordinate_tensor(env):
    """The kernel map is cached by the coordinate storage; a different set of the same size must not hit it."""
    ME, _, O, coord, nbr = env
    conv = ME.MinkowskiConvolution(8, 8, kernel_size=3, stride=1, bias=False, dimension=3).cuda()
    x = torch.randn(len(coord), 8)
    st, xyz = _sparse(ME, coord, x.cuda())
    ref = conv(st).F
    shifted = coord.copy()
    shifted[:, 2] = shifted[::-1, 2]                                      # same size, other geometry (may repeat rows? keep unique)
    shifted = np.unique(shifted, axis=0).astype(np.int32)
    st2, _ = _sparse(ME, shifted, torch.randn(len(shifted), 8).cuda())
    nbr2 = torch.from_numpy(O.nbr27(shifted).astype(np.int64))
    want = O.conv27(st2.F.cpu(), nbr2, conv.kernel.detach().cpu(), None)
    np.testing.assert_allclose(conv(st2).F.detach().cpu().numpy(), want.numpy(), **TOL)
    xyz.add_(0)                                                           # in-place touch bumps the version: cache miss, same answer
    st3, _ = _sparse(ME, coord, x.cuda())
    assert torch.equal(conv(st3).F, ref)


def test_torchac_surface_round_trip(env):
    """BinaryArithmeticCoding (models/module_utils.py:8-40): cdf = [0, 1-p, 1] float32 CPU, symbols int16 CPU."""
    _, torchac, O, coord, nbr = env
    g = torch.Generator().manual_seed(9)
    p = torch.rand(5000, generator=g).clamp(1e-4, 1 - 1e-4)
    sym = (torch.rand(5000, generator=g) < p).to(torch.int16)
    cdf = torch.stack([torch.zeros_like(p), 1 - p, torch.ones_like(p)], dim=1)
    data = torchac.encode_float_cdf(cdf, sym)
    assert isinstance(data, bytes) and len(data) < 5000
    back = torchac.decode_float_cdf(cdf, data)
    assert torch.equal(torch.as_tensor(back).to(torch.int16), sym)
