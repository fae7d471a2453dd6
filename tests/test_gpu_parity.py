"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures recorded from the
reference's own Python.  Integer / byte / index results bit-exact; fp32 results within 1e-4 relative
(BASELINE.json north_star), tolerance written at each assert."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # north_star: logits and losses within 1e-4 relative (fp32)


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import linr_pcgc_b200 as pkg
    from linr_pcgc_b200 import _lib, codec, frame, net, rc, synth, trainer
    _lib.load()  # raises if the extension is missing: there is no fallback

    class NS:
        pass

    ns = NS()
    ns.frame, ns.net, ns.codec, ns.rc, ns.synth, ns.lib, ns.trainer = frame, net, codec, rc, synth, _lib, trainer
    return ns


@pytest.fixture(scope="module")
def O():
    from oracle import linr_oracle
    return linr_oracle


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def _cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _dense_from_compact(anchor, mask, n):
    """Expand (anchor [9,ld], mask [n]) to the dense [n,27] table: present rows of a column are consecutive."""
    anchor = anchor.cpu().numpy()
    mask = mask.cpu().numpy().astype(np.uint32)
    out = -np.ones((n, 27), dtype=np.int32)
    for c in range(9):
        run = np.zeros(n, dtype=np.int32)
        for j in range(3):
            bit = (mask >> np.uint32(3 * c + j)) & 1
            out[:, c + 9 * j] = np.where(bit == 1, anchor[c, :n] + run, -1)
            run += bit.astype(np.int32)
    return out


# ------------------------------------------------------------------------------------------------ coordinate stage
@pytest.mark.parametrize("name", ["tiny", "tiny_s3", "ragged", "mid"])
def test_prepare_frame_bit_exact(L, O, name):
    g = _load(f"int_{name}.npz")
    scale_num = int(g["scale_num"]) if name == "tiny_s3" else None
    fr = L.frame.prepare_frame(_cuda(g["points"]), scale_num, int(g["min_point_num"]), dense=True)
    assert fr.n_scales == int(g["n_scales"])
    assert fr.point_num == int(g["point_num"])
    assert (fr.coord_min == g["coord_min"]).all()
    assert (fr.xyz.cpu().numpy() == g["ori"]).all()
    for s in range(fr.n_scales):
        a, b = fr.scale_off[s], fr.scale_off[s + 1]
        coord = fr.scale_coords(s).cpu().numpy()
        assert (coord == g[f"s{s}_coord"]).all()
        occ = fr.scale_occ(s).cpu().numpy()
        ref_occ = (g[f"s{s}_occ"].astype(np.uint32) << np.arange(8, dtype=np.uint32)).sum(axis=1).astype(np.uint8)
        assert (occ == ref_occ).all()
        nb7 = fr.scale_nbr7(s).cpu().numpy()
        ref7 = (g[f"s{s}_nbr7"].astype(np.uint32) << np.arange(7, dtype=np.uint32)).sum(axis=1).astype(np.uint8)
        assert (nb7 == ref7).all()
        # kernel map: dense table vs oracle (rows are local to the scale in the oracle, global in the frame tables)
        ref27 = O.nbr27(coord)
        got27 = fr.tables.nbr27[a:b].cpu().numpy()
        assert (np.where(got27 >= 0, got27 - a, -1) == ref27).all()
        # upper_layer round trip (datautils/custom_dataset.py:295)
        up = L.frame.octree_up(fr.scale_coords(s), fr.scale_occ(s), fr.bits)
        assert (up.cpu().numpy() == g[f"s{s}_up"]).all()
    n = fr.tables.n_rows
    assert (_dense_from_compact(fr.tables.anchor, fr.tables.mask, n) == fr.tables.nbr27.cpu().numpy()).all()


def test_sort_unique_lookup_bit_exact(L, O):
    g = _load("int_sort.npz")
    xyz = g["xyz"].astype(np.int64)
    lo = xyz.min()
    shifted = _cuda((xyz - lo).astype(np.int32))  # the ABI takes non-negative coordinates
    srt = L.frame.sort_rows(shifted, 8).cpu().numpy() + lo
    assert (srt == g["sorted"]).all()
    uq = L.frame.sort_unique(shifted, 8).cpu().numpy() + lo
    assert (uq == g["uniq"]).all()
    # quantize(.,2) = unique(floor(x/2)) (models/quantize_functions.py:19-30): floor on the original values
    q2 = L.frame.sort_unique(_cuda(((xyz >> 1) - (lo >> 1)).astype(np.int32)), 8).cpu().numpy() + (lo >> 1)
    assert (q2 == g["quant2"]).all()
    # QuickSearchCoord.search / search_coord_idx (models/module_utils.py:260-318) via the hash
    uq_t = _cuda((g["uniq"].astype(np.int64) - lo + 1).astype(np.int32))
    sc = torch.zeros(len(uq_t), dtype=torch.uint8, device="cuda")
    t = L.frame.build_tables(uq_t, sc, keep_hash=True)
    for qname, want_idx in (("query", None), ("query_idx_clamped", g["idx"])):
        q = _cuda((g[qname].astype(np.int64) - lo + 1).astype(np.int32))
        rows = L.frame.hash_lookup(t, q)
        if want_idx is None:
            assert ((rows.cpu().numpy() >= 0).astype(np.uint8) == g["hit"]).all()
        else:
            assert (rows.cpu().numpy() == want_idx).all()


def test_empty_and_single_point(L):
    one = torch.tensor([[5, 6, 7]], dtype=torch.int32, device="cuda")
    fr = L.frame.prepare_frame(one, None, 64)
    assert fr.n_scales == 1 and fr.point_num == 1
    assert fr.scale_occ(0).cpu().tolist() == [1]  # after min subtraction the point is (0,0,0): octant 0
    up = L.frame.octree_up(fr.scale_coords(0), fr.scale_occ(0), 4)
    assert up.cpu().tolist() == [[0, 0, 0]]
    empty = torch.zeros((0, 3), dtype=torch.int32, device="cuda")
    assert L.frame.sort_unique(empty, 4).shape[0] == 0
    with pytest.raises(L.lib.LinrError):
        L.frame.prepare_frame(empty, None, 64)
    with pytest.raises(L.lib.LinrError):
        L.frame.prepare_frame(torch.zeros((4, 3), dtype=torch.int32), None, 64)  # host tensor: no CPU path


def test_c_abi_error_codes_and_messages(L, O):
    """Error behaviour of the C ABI (include/linr_b200.h): negative return code + linr_last_error(), nothing launched.
    The reference signals the same situations with Python asserts / exceptions."""
    import ctypes as C
    lib = L.lib.load()
    g, S, sd, flat, fr = _net_case(L, O)
    params = flat.cuda()
    run = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=True)
    rows = fr.tables.rows()
    stream = L.lib.stream_ptr()
    ptr = L.lib.ptr
    # workspace too small
    rc = lib.linr_net_forward(ptr(params), S, C.byref(rows), 1, 1.0, None, None, ptr(run.bits), ptr(run.ws), 1024, stream)
    assert rc < 0 and b"workspace too small" in lib.linr_last_error()
    rc = lib.linr_net_backward(ptr(params), S, C.byref(rows), ptr(torch.empty_like(params)), ptr(run.ws), 1024, stream)
    assert rc < 0 and b"workspace too small" in lib.linr_last_error()
    # scale_num out of range, stage out of range, bad bit depth, Adam step 0
    assert lib.linr_net_forward(ptr(params), 0, C.byref(rows), 0, 0.0, None, None, None, ptr(run.ws), run.ws.numel(), stream) < 0
    assert b"scale_num" in lib.linr_last_error()
    assert lib.linr_net_decode_stage(ptr(params), S, C.byref(rows), 8, None, ptr(run.cdf), ptr(run.ws), run.ws.numel(), stream) < 0
    assert b"stage" in lib.linr_last_error()
    q = torch.empty(params.numel(), dtype=torch.uint8, device="cuda")
    st = torch.empty(4, device="cuda")
    assert lib.linr_param_quant(ptr(params), params.numel(), 9, ptr(q), ptr(torch.empty_like(params)), ptr(st), stream) < 0
    assert b"bitdepth" in lib.linr_last_error()
    z = torch.zeros_like(params)
    assert lib.linr_adam_fused(ptr(params.clone()), ptr(z), ptr(z.clone()), ptr(z.clone()), params.numel(), 0, 0.01, 0.9, 0.999, 1e-8, 1e-4, stream) < 0
    # unsupported channel counts of the single-layer entry points
    x = torch.zeros(fr.tables.n_rows, 8, device="cuda")
    assert lib.linr_spconv27_fwd(ptr(x), 6, ptr(torch.zeros(27 * 6 * 8, device="cuda")), None, ptr(x.clone()), 8, C.byref(rows), 0, stream) < 0
    assert b"channels" in lib.linr_last_error()
    # the Python wrappers turn every failure into LinrError and refuse host tensors
    with pytest.raises(L.lib.LinrError):
        L.net.spconv27_fwd(x.cpu(), torch.zeros(27, 8, 8), None, fr.tables)
    torch.cuda.synchronize()
    # nothing above left the device in an error state: a normal call still works
    out = run.forward(params, fr.tables, train=True, loss_scale=1.0 / fr.point_num)
    assert float(out["bits"].item()) > 0


@pytest.mark.parametrize("shape,bits", [("loot", 10), ("owlii", 11)])
def test_full_size_octree_round_trip(L, shape, bits):
    """BASELINE.json full sizes: size-independent properties (round trip, sortedness, idempotence)."""
    pts = L.synth.make_sequence(shape, 1, device="cuda")[0]
    perm = torch.randperm(pts.shape[0], device="cuda")
    fr = L.frame.prepare_frame(torch.cat([pts[perm], pts[perm[:1000]]]), None, 64)  # shuffled + duplicates
    assert fr.point_num == pts.shape[0]
    assert (fr.xyz + torch.from_numpy(fr.coord_min).cuda() == pts).all()
    assert fr.n_scales == (7 if shape == "loot" else 8)
    child = fr.xyz
    for s in range(fr.n_scales):
        c = fr.scale_coords(s).to(torch.int64)
        key = (c[:, 0] << 42) | (c[:, 1] << 21) | c[:, 2]
        assert (key[1:] > key[:-1]).all()  # strictly sorted = unique
        up = L.frame.octree_up(fr.scale_coords(s), fr.scale_occ(s), fr.bits)
        assert up.shape == child.shape and (up == child).all()
        again = L.frame.sort_unique(fr.scale_coords(s), fr.bits)
        assert (again == fr.scale_coords(s)).all()
        child = fr.scale_coords(s)
    # kernel map symmetry: pair (i -> o, k) <=> (o -> i, 26-k)
    n = fr.scale_off[1]
    sc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    t = L.frame.build_tables(fr.scale_coords(0).contiguous(), sc, dense=True)
    nb = t.nbr27.to(torch.int64)
    assert (nb[:, 13] == torch.arange(n, device="cuda")).all()
    for k in (0, 5, 12, 22, 26):
        o = torch.nonzero(nb[:, k] >= 0)[:, 0]
        assert (nb[nb[o, k], 26 - k] == o).all()
    assert int(fr.scale_coords(fr.n_scales - 1).max()) < 256  # test_utils.py:221


# ------------------------------------------------------------------------------------------------ network
NET_FIXTURES = ["tiny", "mid"]   # mid: ~40k points, 5 scales, ~14k rows -> several 256-row chunks of partial sums per kernel


def _net_case(L, O, name="tiny"):
    g = _load(f"net_{name}.npz")
    S = int(g["scale_num"])
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}
    flat = O.flatten_params(sd, S).contiguous()
    assert flat.numel() == L.net.param_count(S)
    fr = L.frame.prepare_frame(_cuda(g["points"]), None, 64)
    return g, S, sd, flat, fr


def test_param_layout_matches_checkpoint_contract(L, O):
    for S in (3, 7, 8):
        spec = O.param_spec(S)
        offs = L.net.param_offsets(S)
        want, o = [], 0
        for _, shp in spec:
            want.append(o)
            o += int(np.prod(shp))
        assert offs == want and L.net.param_count(S) == o
    assert L.net.param_count(7) == 54712


@pytest.mark.parametrize("fixture", NET_FIXTURES)
def test_forward_probs_bits_cdf(L, O, fixture):
    g, S, sd, flat, fr = _net_case(L, O, fixture)
    run = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=False)
    out = run.forward(flat.cuda(), fr.tables, want_probs=True, want_cdf=True, want_bits=True)
    probs = out["probs"].cpu().numpy()
    bits = float(out["bits"].item())
    assert abs(bits - float(g["bits"])) <= RTOL * float(g["bits"])
    for s in range(fr.n_scales):
        a, b = fr.scale_off[s], fr.scale_off[s + 1]
        np.testing.assert_allclose(probs[:, a:b].T, g[f"s{s}_probs"], rtol=RTOL, atol=1e-6)
    # 16-bit CDF boundary = torchac's conversion of the GPU's own probability, bit-exact
    cdf = out["cdf"].cpu().numpy().view(np.uint16)
    assert (cdf == O.cdf_u16_binary(probs.reshape(-1)).reshape(cdf.shape)).all()


@pytest.mark.parametrize("fixture", NET_FIXTURES)
def test_backward_adam_match_and_deterministic(L, O, fixture):
    g, S, sd, flat, fr = _net_case(L, O, fixture)
    P = flat.numel()
    run = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=True)
    params = flat.cuda()
    grads = []
    for _ in range(2):
        grad = torch.full((P,), float("nan"), device="cuda")
        out = run.forward(params, fr.tables, train=True, loss_scale=1.0 / fr.point_num)
        run.backward(params, fr.tables, grad)
        grads.append(grad.cpu().numpy())
    assert np.array_equal(grads[0], grads[1])  # run-to-run bitwise reproducible (no fp atomics)
    ref = g["grad_flat"]
    assert np.isfinite(grads[0]).all()
    # per tensor: max abs error relative to the tensor's own scale
    offs = L.net.param_offsets(S) + [P]
    names = [n for n, _ in O.param_spec(S)]
    for i, n in enumerate(names):
        a, b = offs[i], offs[i + 1]
        scale = max(np.abs(ref[a:b]).max(), 1e-6 * np.abs(ref).max())
        err = np.abs(grads[0][a:b] - ref[a:b]).max() / scale
        assert err < 5e-4, (n, err)
    assert np.abs(grads[0] - ref).max() / np.abs(ref).max() < RTOL
    # Adam (main.py:231-237), one step from zero moments, fed the recorded gradient
    m = torch.zeros(P, device="cuda")
    v = torch.zeros(P, device="cuda")
    p2 = params.clone()
    L.net.adam_step(p2, _cuda(ref), m, v, 1, 0.01)
    np.testing.assert_allclose(p2.cpu().numpy(), g["flat_after_adam"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("fixture", NET_FIXTURES)
def test_param_quant_bit_exact(L, O, fixture):
    g, S, sd, flat, fr = _net_case(L, O, fixture)
    q, recon, stats = L.net.param_quant(flat.cuda(), 8)
    np.testing.assert_array_equal(recon.cpu().numpy(), g["q_recon"])
    st = stats.cpu().numpy()
    assert st[0] == g["q_min"] and st[1] == g["q_max"] and st[2] == g["q_mu"] and st[3] == g["q_b"]
    qo, _, _, _ = O.quant_uniform2(flat, 8)
    np.testing.assert_array_equal(q.cpu().numpy(), qo.numpy().astype(np.uint8))


def test_model_coder_versions_and_bit_depths(L, O):
    """(1) cdf_version 2 fixes the Laplace row quirk ([0,c1..cL] instead of [c1..cL,0],
    model_size_est.py:473-478): same symbols, decodable, never larger; (2) model_bitdepth 9..16 -- which the reference
    writes but cannot read back (model_size_est.py:546-548) -- round-trips through 16-bit symbols."""
    from linr_pcgc_b200 import model_compression as MC
    g, S, sd, flat, fr = _net_case(L, O, "mid")
    params = flat.cuda()
    n = params.numel()
    sizes = {}
    for ver in (MC.CDF_REFERENCE, MC.CDF_FIXED):
        c = MC.compress_model(params, 8, ver)
        assert c["enc_mode"] == 2 and c["cdf_version"] == ver
        rec = MC.decompress_model(dict(c), n)
        assert torch.equal(rec, c["recon_ret"])
        sizes[ver] = len(c["final_bytes"])
    assert sizes[MC.CDF_FIXED] <= sizes[MC.CDF_REFERENCE]
    assert MC.compress_model(params, 8)["final_bytes"] == g["q_bytes"].tobytes()       # default = the reference's bitstream
    rng = float(params.max() - params.min())
    for bd in (9, 12, 16):
        c = MC.compress_model(params, bd)
        assert c["enc_mode"] in (0, 1) and c["bitdepth"] == bd
        rec = MC.decompress_model(dict(c), n)
        assert torch.equal(rec, c["recon_ret"])
        assert float((rec - params).abs().max()) <= 0.51 * rng / (2 ** bd - 1) + 1e-6   # uniform quantiser: half a step (+ fp32 rounding)
        q = (c["quant"].to(torch.int32) & 0xFFFF)
        assert int(q.max()) == 2 ** bd - 1 and int(q.min()) == 0
    print("model.bin bytes: reference row %d, fixed row %d (%.2f %% smaller)" % (sizes[1], sizes[2], 100 * (1 - sizes[2] / sizes[1])))


def test_single_layer_conv_entry_points(L, O):
    g = _load("int_mid.npz")
    coord = g["s0_coord"]
    n = len(coord)
    t = L.frame.build_tables(_cuda(coord), torch.zeros(n, dtype=torch.uint8, device="cuda"), dense=True)
    nbr = t.nbr27.cpu().to(torch.int64)
    gen = torch.Generator().manual_seed(5)
    for cin, cout in ((8, 8), (8, 4), (4, 4), (4, 8)):
        x = torch.randn(n, cin, generator=gen)
        W = torch.randn(27, cin, cout, generator=gen) * 0.2
        b = torch.randn(cout, generator=gen)
        ref = O.conv27(x, nbr, W, b)
        y = L.net.spconv27_fwd(x.cuda(), W.cuda(), b.cuda(), t)
        np.testing.assert_allclose(y.cpu().numpy(), ref.numpy(), rtol=RTOL, atol=1e-5)
        # autograd of the oracle conv gives the reference gradients
        xr, Wr = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
        dy = torch.randn(n, cout, generator=gen)
        O.conv27(xr, nbr, Wr, b).backward(dy)
        dx = L.net.spconv27_bwd_in(dy.cuda(), W.cuda(), t)
        np.testing.assert_allclose(dx.cpu().numpy(), xr.grad.numpy(), rtol=RTOL, atol=1e-5)
        if (cin, cout) != (4, 8):
            dW, db = L.net.spconv27_bwd_w(x.cuda(), dy.cuda(), t)
            sc = Wr.grad.abs().max().item()
            assert (dW.cpu() - Wr.grad).abs().max().item() / sc < RTOL
            np.testing.assert_allclose(db.cpu().numpy(), dy.sum(0).numpy(), rtol=RTOL, atol=1e-4)


def test_full_size_conv_identities(L):
    """Size-independent properties of the three conv kernels on a loot-sized frame (all 7 scales, 277 k rows), no
    oracle needed: (1) an all-ones input gives y[o] = sum over the PRESENT offsets of the column sums of W[k] -- an
    independent formula on the kernel map's presence bits; (2) adjointness <conv(x), dy> = <x, conv^T(dy)> ties the
    grad-input kernel to the forward one; (3) <conv(x), dy> = <W, dW> and db = sum(dy) tie the weight-gradient kernel to
    it; (4) linearity."""
    pts = L.synth.make_sequence("loot", 1, device="cuda")[0]
    fr = L.frame.prepare_frame(pts, None, 64)
    t = fr.tables
    n = t.n_rows
    assert n > 250_000
    gen = torch.Generator(device="cuda").manual_seed(17)
    bits = torch.stack([(t.mask.to(torch.int64) >> (3 * (k % 9) + k // 9)) & 1 for k in range(27)], dim=1).double()   # [n,27]
    for cin, cout in ((8, 8), (8, 4), (4, 4)):
        W = torch.randn(27, cin, cout, generator=gen, device="cuda") * 0.2
        x = torch.randn(n, cin, generator=gen, device="cuda")
        dy = torch.randn(n, cout, generator=gen, device="cuda")
        # (1) ones
        y1 = L.net.spconv27_fwd(torch.ones(n, cin, device="cuda"), W, None, t)
        want = bits @ W.double().sum(dim=1)                                   # [n,27] @ [27,cout]
        assert (y1.double() - want).abs().max().item() < 1e-4
        # (2) adjoint of the forward conv
        y = L.net.spconv27_fwd(x, W, None, t)
        dx = L.net.spconv27_bwd_in(dy, W, t)
        lhs = (y.double() * dy.double()).sum().item()
        rhs = (x.double() * dx.double()).sum().item()
        scale = (y.double().abs() * dy.double().abs()).sum().item()
        assert abs(lhs - rhs) <= 1e-6 * scale, (cin, cout, lhs, rhs)
        # (3) weight gradient
        dW, db = L.net.spconv27_bwd_w(x, dy, t)
        assert abs(lhs - (W.double() * dW.double()).sum().item()) <= 1e-6 * scale
        ref_db = dy.double().sum(0)
        assert (db.double() - ref_db).abs().max().item() <= 1e-5 * dy.double().abs().sum(0).max().item()
        # (4) linearity in the input
        x2 = torch.randn(n, cin, generator=gen, device="cuda")
        ya = L.net.spconv27_fwd(2.0 * x - 0.5 * x2, W, None, t)
        yb = 2.0 * y - 0.5 * L.net.spconv27_fwd(x2, W, None, t)
        assert (ya - yb).abs().max().item() < 1e-4 * max(1.0, yb.abs().max().item())


# ------------------------------------------------------------------------------------------------ coding
@pytest.mark.parametrize("fixture", NET_FIXTURES)
def test_encode_decode_lossless_and_bpp(L, O, fixture):
    g, S, sd, flat, fr = _net_case(L, O, fixture)
    params = flat.cuda()
    run = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=False)
    all_bytes = L.codec.encode_frame(run, params, fr)
    tot = sum(len(b) for b in all_bytes) * 8
    # bpp within 0.5 % of the reference flow's bitstream (north_star)
    assert abs(tot - int(g["all_bit"])) <= 0.005 * int(g["all_bit"])
    for s, b in enumerate(all_bytes):
        assert abs(len(b) - len(g[f"s{s}_bytes"])) <= max(2, 0.005 * len(g[f"s{s}_bytes"]))
    low = fr.scale_coords(fr.n_scales - 1).contiguous()
    dec = L.codec.decode_frame(run, params, all_bytes, low)
    assert dec.shape == fr.xyz.shape and (dec == fr.xyz).all()  # decoder.py:140
    assert (dec.cpu().numpy() == g["dec_coord"]).all()
    # the product's host coder is bitstream-compatible with the oracle's torchac restatement
    from oracle import rc as orc
    cdf, occ, R = L.codec.frame_cdfs_to_host(run, params, fr)
    a, b = fr.scale_off[0], fr.scale_off[1]
    for k in (0, 7):
        sym = ((occ[a:b] >> k) & 1).astype(np.int16)
        n = b - a
        rows = np.stack([np.zeros(n, np.uint16), cdf[k, a:b], np.zeros(n, np.uint16)], axis=1)
        want = orc.encode_u16(rows, sym)
        got = L.codec.unpack_bitstream(all_bytes[0])[k]
        assert got == want
        assert (orc.decode_u16(rows, got, n) == sym).all()


def test_weight_bank_variant_is_bit_identical(L, O):
    """The training forward reads the conv weights from the constant bank (uniform-register FFMA2 operands), the
    coding forward from shared memory: same arithmetic, same order -> the same probability bits."""
    g, S, sd, flat, fr = _net_case(L, O)
    params = flat.cuda()
    tr_run = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=True)
    inf_run = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=False)
    a = tr_run.forward(params, fr.tables, train=True, loss_scale=1.0 / fr.point_num, want_probs=True, want_cdf=True)
    pa, ca, ba = a["probs"].clone(), a["cdf"].clone(), a["bits"].clone()
    b = inf_run.forward(params, fr.tables, train=False, want_probs=True, want_cdf=True)
    assert torch.equal(pa, b["probs"]) and torch.equal(ca, b["cdf"]) and torch.equal(ba, b["bits"])
    # and on a side stream (which does not own the bank: shared-memory variant even with train=True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run2 = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=True)
        c = run2.forward(params, fr.tables, train=True, loss_scale=1.0 / fr.point_num, want_probs=True)
        grad2 = torch.empty_like(params)
        run2.backward(params, fr.tables, grad2)
        pc = c["probs"].clone()
    side.synchronize()
    grad1 = torch.empty_like(params)
    tr_run.backward(params, fr.tables, grad1)
    torch.cuda.synchronize()
    assert torch.equal(pc, pa) and torch.equal(grad1, grad2)


def test_contexts_take_turns_at_the_weight_bank(L, O):
    """linr_ctx (include/linr_b200.h): a trainer on stream A, destroyed, then a trainer on stream B -- and two live
    trainers stepping alternately -- all run their training calls on the constant-bank kernels (round 1: the first
    (thread, stream) pair kept the bank for the life of the process and everybody else fell back silently)."""
    g, S, sd, flat, fr = _net_case(L, O)
    params = flat.cuda()
    grads = []

    def one_iteration(run):
        run.forward(params, fr.tables, train=True, loss_scale=1.0 / fr.point_num)
        grad = torch.empty_like(params)
        run.backward(params, fr.tables, grad)
        return grad

    a = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=True)
    grads.append(one_iteration(a))
    assert a.bank_calls() == 2 and a.bank_launches() > 10           # forward + backward, every multi-group conv launch
    a.close()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        b = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=True)
        grads.append(one_iteration(b))
        side.synchronize()
        assert b.bank_calls() == 2 and b.bank_launches() > 10
    c = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=True)    # two live trainers, one host thread: both get their turn
    for _ in range(2):
        with torch.cuda.stream(side):
            grads.append(one_iteration(b))
        grads.append(one_iteration(c))
    torch.cuda.synchronize()
    assert b.bank_calls() == 6 and c.bank_calls() == 4
    assert all(torch.equal(grads[0], x) for x in grads[1:])          # same bits whoever held the bank, whatever the stream
    inf = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=False)
    assert inf.bank_calls() == 0 and inf.ctx is None                 # coding runners never touch the bank
    b.close(), c.close()


def test_second_stream_changes_no_bit(L, O):
    """linr_side_stream_enable (include/linr_b200.h): the weight-gradient launches of a backward call and the bit-input
    ConvA of a training forward run on the context's second stream, forked from / joined to the caller's stream inside
    the call.  Same kernels, same partial sums: bits, gradient and the parameters after several optimiser steps on
    frames that reuse one workspace are bitwise those of the one-stream schedule -- also when the caller's stream is
    not the default stream and other work is queued behind the call."""
    lib = L.lib.load()
    pts = L.synth.make_sequence("tiny", 3)
    frames = [L.frame.prepare_frame(p.cuda(), None, 64) for p in pts]
    S = frames[0].n_scales
    mr = max(f.tables.n_rows for f in frames)

    def run(on, stream=None):
        prev = lib.linr_side_stream_enable(1 if on else 0)
        try:
            with torch.cuda.stream(stream or torch.cuda.current_stream()):
                tr = L.trainer.GopTrainer(S, "cuda", seed=11, max_rows=mr)
                bits, grads = [], []
                for it in range(6):
                    fr = frames[it % 3]
                    out = tr.runner.forward(tr.state.params, fr.tables, train=True, loss_scale=1.0 / fr.point_num)
                    tr.runner.backward(tr.state.params, fr.tables, tr.grad)
                    bits.append(out["bits"].clone()), grads.append(tr.grad.clone())   # queued right behind the call
                    tr.step(fr)
                res = (torch.stack(bits), torch.stack(grads), tr.state.params.clone())
                tr.runner.close()
            torch.cuda.synchronize()
            return res
        finally:
            lib.linr_side_stream_enable(prev)

    base = run(False)
    other = torch.cuda.Stream()
    other.wait_stream(torch.cuda.current_stream())
    for res in (run(True), run(True, other), run(True)):
        for a, b in zip(base, res):
            assert torch.equal(a, b)


def test_two_trainers_in_two_host_threads_match_serial_runs(L, O):
    """The constant weight bank is owned by one (host thread, stream) pair; a second trainer running concurrently in
    another host thread -- even on the same stream -- must fall back to the shared-memory kernels and still produce
    exactly the parameters of a serial run."""
    import threading
    pts = L.synth.make_sequence("tiny", 2)
    frames = [L.frame.prepare_frame(p.cuda(), None, 64) for p in pts]
    S = frames[0].n_scales

    def train(seed, out, key):
        torch.cuda.set_device(0)
        tr = L.trainer.GopTrainer(S, "cuda", seed=seed, max_rows=max(f.tables.n_rows for f in frames))
        tr.fit(frames, 3)
        torch.cuda.synchronize()
        out[key] = tr.state.params.clone()

    serial, conc = {}, {}
    train(1, serial, "a")
    train(2, serial, "b")
    ths = [threading.Thread(target=train, args=(1, conc, "a")), threading.Thread(target=train, args=(2, conc, "b"))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert torch.equal(serial["a"], conc["a"]) and torch.equal(serial["b"], conc["b"])


def test_batched_and_sequential_cdfs_identical(L, O):
    """Encoder (teacher-forced, all scales in one launch set) and decoder (per scale, stage by stage) must see
    bit-identical 16-bit CDFs: the precondition of lossless decoding."""
    g, S, sd, flat, fr = _net_case(L, O)
    params = flat.cuda()
    run = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=False)
    enc = run.forward(params, fr.tables, want_cdf=True, want_probs=True, want_bits=False)
    enc_cdf = enc["cdf"].clone()
    enc_p = enc["probs"].clone()
    run2 = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=False)
    for s in range(fr.n_scales):
        a, b = fr.scale_off[s], fr.scale_off[s + 1]
        n = b - a
        t = L.frame.build_tables(fr.scale_coords(s).contiguous(), torch.full((n,), s, dtype=torch.uint8, device="cuda"),
                                 fr.scale_occ(s).contiguous())
        run2.decode_begin(params, t)
        for k in range(8):
            cdf_k, p_k = run2.decode_stage(params, t, k, want_probs=True)
            assert torch.equal(cdf_k, enc_cdf[k, a:b])
            assert torch.equal(p_k, enc_p[k, a:b])


def test_rc_edges(L):
    rng = np.random.default_rng(11)
    for n in (0, 1, 7, 4097):
        cdf = rng.integers(1, 65536, size=n).astype(np.uint16)
        cdf[: n // 8] = 1
        cdf[n // 8: n // 4] = 65535
        sym = rng.integers(0, 2, size=n).astype(np.uint8)
        b = L.rc.encode_binary(cdf, sym)
        assert (L.rc.decode_binary(cdf, b, n) == sym).all()
    packed = rng.integers(0, 256, size=1000).astype(np.uint8)
    cdfs = [rng.integers(1, 65536, size=1000).astype(np.uint16) for _ in range(8)]
    outs = L.rc.encode_binary_batch(cdfs, [packed] * 8, list(range(8)), threads=4)
    for k in range(8):
        assert outs[k] == L.rc.encode_binary(cdfs[k], (packed >> k) & 1)


def test_full_size_encode_decode_lossless(L, O):
    """loot-sized frame (~780k points, 7 scales, 2.3 M symbols): encode -> decode round trip, random-init model."""
    pts = L.synth.make_sequence("loot", 1, device="cuda")[0]
    fr = L.frame.prepare_frame(pts, None, 64)
    S = fr.n_scales
    flat = O.flatten_params(O.init_params(S, seed=3), S).cuda()
    run = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=False)
    all_bytes = L.codec.encode_frame(run, flat, fr)
    dec = L.codec.decode_frame(run, flat, all_bytes, fr.scale_coords(S - 1).contiguous())
    assert dec.shape == fr.xyz.shape and (dec == fr.xyz).all()


def test_gop_overfit_tracks_oracle_training_and_bpp(L, O):
    """The north star's end criterion on a small GOP: the same overfitting schedule (one Adam step per frame,
    StepLR per step, main.py:297-322) run by the CPU oracle and by the CUDA path from the same initial parameters gives
    the same loss curve (1e-3 relative after 8 optimiser steps) and bitstreams within 0.5 % of each other in size."""
    import torch
    from linr_pcgc_b200 import params as P
    pts = L.synth.make_sequence("tiny", 2)
    frames = [L.frame.prepare_frame(p.cuda(), None, 64) for p in pts]
    S = frames[0].n_scales
    flat0 = P.init_flat(S, seed=13)
    # ---- oracle: 4 epochs x 2 frames
    ofr = [O.prepare_frame(p.numpy(), S, 64) for p in pts]
    nbrs = [[torch.from_numpy(O.nbr27(sc["coord"]).astype(np.int64)) for sc in f["scales"]] for f in ofr]
    flat = flat0.clone()
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    ref_losses, step = [], 0
    for ep in range(4):
        acc = []
        for f, nb in zip(ofr, nbrs):
            p = flat.clone().requires_grad_(True)
            bits = O.frame_bits(O.unflatten_params(p, S), f, nb)
            (bits / f["point_num"]).backward()
            step += 1
            with torch.no_grad():
                O.adam_step_reference([flat], [p.grad], [m], [v], step=step, lr=O.lr_at_step(step - 1))
            acc.append(float(bits.detach()) / f["point_num"])
        ref_losses.append(float(np.mean(acc)))
    with torch.no_grad():
        ref_bytes = sum(len(b) for f in ofr for b in O.encode_frame(O.unflatten_params(flat, S), f))
    # ---- CUDA path
    tr = L.trainer.GopTrainer(S, "cuda", max_rows=max(f.tables.n_rows for f in frames),
                              state=L.trainer.OptimState(flat0.clone().cuda(), torch.zeros_like(flat0).cuda(),
                                                         torch.zeros_like(flat0).cuda(), 0, 0, 0.01))
    losses = tr.fit(frames, 4)
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-3)
    assert losses[-1] < losses[0]
    run = L.net.NetRunner(S, max(f.tables.n_rows for f in frames), "cuda", train=False)
    got_bytes = sum(len(b) for f in frames for b in L.codec.encode_frame(run, tr.state.params, f))
    assert abs(got_bytes - ref_bytes) <= 0.005 * ref_bytes, (got_bytes, ref_bytes)
    np.testing.assert_allclose(tr.state.params.cpu().numpy(), flat.numpy(), rtol=0, atol=2e-3)


def test_two_gops_track_torch_adam_steplr_oracle(L, O):
    """GOP-level schedule semantics against the classes the reference itself uses: the CPU oracle is stepped by
    torch.optim.Adam (L2 1e-4) + StepLR per frame with the per-epoch min_lr floor, and a second GOP loads the first one's
    optimizer state dict and builds a fresh StepLR (main.py:231-252,319-321,433-437,102-104).  44 optimiser steps cross
    eight StepLR boundaries, the floor, and a GOP boundary whose step count is not a multiple of step_size.  The CUDA
    path (GopTrainer, seeded by OptimState) must use the same lr at every step and track the loss curve."""
    import torch
    from linr_pcgc_b200 import params as P
    lr0, gamma, step_size, min_lr = 0.01, 0.7, 5, 4e-3
    pts = L.synth.make_sequence("tiny", 8)
    gops = [pts[:4], pts[4:]]
    epochs = [6, 5]
    S = L.frame.prepare_frame(pts[0].cuda(), None, 64).n_scales
    flat0 = P.init_flat(S, seed=21)
    # ---- oracle
    w = torch.nn.Parameter(flat0.clone())
    ref_losses, ref_lrs, sd = [], [], None
    for gp, ep_n in zip(gops, epochs):
        ofr = [O.prepare_frame(p.numpy(), S, 64) for p in gp]
        nbrs = [[torch.from_numpy(O.nbr27(sc["coord"]).astype(np.int64)) for sc in f["scales"]] for f in ofr]
        opt = torch.optim.Adam([w], lr=lr0, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
        if sd is not None:
            opt.load_state_dict(sd)
        sch = torch.optim.lr_scheduler.StepLR(opt, step_size=step_size, gamma=gamma)
        for _ in range(ep_n):
            acc = []
            for f, nb in zip(ofr, nbrs):
                opt.zero_grad()
                bits = O.frame_bits(O.unflatten_params(w, S), f, nb)
                (bits / f["point_num"]).backward()
                ref_lrs.append(opt.param_groups[0]["lr"])
                opt.step()
                sch.step()
                acc.append(float(bits.detach()) / f["point_num"])
            for g in opt.param_groups:
                if g["lr"] < min_lr:
                    g["lr"] = min_lr
            ref_losses.append(float(np.mean(acc)))
        sd = opt.state_dict()
    # ---- CUDA path
    state = L.trainer.OptimState(flat0.clone().cuda(), torch.zeros_like(flat0).cuda(), torch.zeros_like(flat0).cuda(), 0, 0, lr0)
    losses, lrs = [], []
    for gp, ep_n in zip(gops, epochs):
        frames = [L.frame.prepare_frame(p.cuda(), S, 64) for p in gp]
        tr = L.trainer.GopTrainer(S, "cuda", lr0, gamma, step_size, min_lr, max_rows=max(f.tables.n_rows for f in frames), state=state)
        for _ in range(ep_n):
            # record the lr each step will use, then run the epoch
            st = tr.state
            probe = L.trainer.OptimState(st.params, st.m, st.v, st.step, st.sched_step, st.lr)
            for _ in frames:
                lrs.append(probe.lr)
                L.trainer.sched_after_step(probe, step_size, gamma)
            losses += tr.fit(frames, 1)
        state = tr.state
    assert len(ref_lrs) == 44 and len(set(np.round(ref_lrs, 12))) >= 5
    np.testing.assert_allclose(lrs, ref_lrs, rtol=1e-12, atol=0)
    assert tr.state.lr == pytest.approx(sd["param_groups"][0]["lr"], rel=1e-12) and tr.state.step == 44
    np.testing.assert_allclose(losses, ref_losses, rtol=2e-3)
    # parameters: Adam's m / sqrt(v) is scale-free, so the few entries whose gradient sits at rounding-noise level drift by
    # up to lr per step between two fp32 implementations; compared in norm, with a bound on how many entries stray
    got, want = tr.state.params.cpu().numpy(), w.detach().numpy()
    assert np.linalg.norm(got - want) <= 2e-2 * np.linalg.norm(want)
    assert (np.abs(got - want) > 5e-3).mean() < 5e-3
    m_ref = sd["state"][0]["exp_avg"].numpy()
    assert np.linalg.norm(tr.state.m.cpu().numpy() - m_ref) <= 0.15 * np.linalg.norm(m_ref)   # first moments of the noise-level entries differ most


def test_owlii_sized_iteration_and_codec(L, O):
    """Owlii-shaped frame (11-bit, ~2.5 M points, 8 scales, BASELINE.json configs[3]): two frame-iterations are
    bitwise reproducible run to run (no floating-point atomics), the loss is finite, and the trained model codes the
    frame losslessly at that size."""
    import torch
    pts = L.synth.make_sequence("owlii", 1, device="cuda")[0]
    fr = L.frame.prepare_frame(pts, None, 64)
    S = fr.n_scales
    assert S == 8 and fr.point_num > 2_300_000
    outs = []
    for _ in range(2):
        tr = L.trainer.GopTrainer(S, "cuda", seed=11, max_rows=fr.tables.n_rows)
        b0 = float(tr.step(fr).item())
        b1 = float(tr.step(fr).item())
        assert b0 == b0 and b1 == b1 and 0 < b1 < 8.0 * 8 * fr.tables.n_rows
        outs.append((b0, b1, tr.state.params.clone(), tr.grad.clone()))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert torch.equal(outs[0][2], outs[1][2]) and torch.equal(outs[0][3], outs[1][3])
    run = L.net.NetRunner(S, fr.tables.n_rows, "cuda", train=False)
    all_bytes = L.codec.encode_frame(run, outs[0][2], fr)
    dec = L.codec.decode_frame(run, outs[0][2], all_bytes, fr.scale_coords(S - 1).contiguous())
    assert dec.shape == fr.xyz.shape and (dec == fr.xyz).all()


def test_concurrent_decode_mvub_sized(L, O):
    """MVUB-shaped frames (~300k points, BASELINE.json configs[4]): pipelined encode of several frames, then the
    frames decoded concurrently on separate streams/threads; bytes identical to the frame-by-frame encoder."""
    pts = L.synth.make_sequence("mvub10", 3, device="cuda")
    frames = [L.frame.prepare_frame(p, None, 64) for p in pts]
    S = frames[0].n_scales
    flat = O.flatten_params(O.init_params(S, seed=4), S).cuda()
    run = L.net.NetRunner(S, max(f.tables.n_rows for f in frames), "cuda", train=False)
    enc = L.codec.encode_frames(run, flat, frames, threads=4)
    for f, e in zip(frames, enc):
        assert e == L.codec.encode_frame(run, flat, f, threads=2)
    jobs = [(e, f.scale_coords(f.n_scales - 1).contiguous()) for f, e in zip(frames, enc)]
    for workers in (1, 3):
        dec = L.codec.decode_frames(flat, S, jobs, workers=workers)
        for d, f in zip(dec, frames):
            assert d.shape == f.xyz.shape and (d == f.xyz).all()


def test_batched_lockstep_decoder_is_lossless_and_matches_per_frame_decoder(L, O):
    """codec.decode_frames_batched: all frames of a batch decoded in lockstep (concatenated parents, one launch set and
    one round trip per stage for the whole batch).  Same points as the per-frame decoder, with ragged scale counts
    (a small frame stops after fewer scales), more frames than one batch, and a batch of one."""
    pts = L.synth.make_sequence("plumbing", 5, device="cuda")
    small = L.synth.make_sequence("tiny", 1, device="cuda")[0]
    frames = [L.frame.prepare_frame(p, None, 64) for p in pts]
    S = frames[0].n_scales
    frames.insert(2, L.frame.prepare_frame(small, S, 64))        # fewer scales than the others
    assert frames[2].n_scales < S
    flat = O.flatten_params(O.init_params(S, seed=9), S).cuda()
    run = L.net.NetRunner(S, max(f.tables.n_rows for f in frames), "cuda", train=False)
    enc = L.codec.encode_frames(run, flat, frames, threads=4)
    jobs = [(e, f.scale_coords(f.n_scales - 1).contiguous()) for f, e in zip(frames, enc)]
    ref = L.codec.decode_frames(flat, S, jobs, workers=2)
    for max_batch in (16, 4, 1):
        dec = L.codec.decode_frames_batched(flat, S, jobs, max_batch=max_batch, workers=2 if max_batch > 1 else 1, threads=4)
        for d, r, f in zip(dec, ref, frames):
            assert d.dtype == torch.int32 and d.shape == f.xyz.shape and torch.equal(d, r) and torch.equal(d, f.xyz)
