"""CPU-side tests of the product's host logic (no GPU): the C-ABI library loads and exports every declared symbol,
the host range coder is bitstream-identical to the oracle's torchac restatement, the parameter contract matches the
reference checkpoint, file formats round-trip, and the multi-GPU plumbing works on world_size 2 over gloo."""
import json
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT

import linr_pcgc_b200  # noqa: F401
from linr_pcgc_b200 import _lib, codec, dist as D, model_compression, params as P, pointio, rc, synth
from linr_pcgc_b200 import main as cli


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "linr_b200.h")).read()
    declared = set(re.findall(r"\b(linr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/linr_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert lib.linr_version() >= 100
    assert lib.linr_param_count(7) == 54712


def test_param_contract_matches_reference_checkpoint():
    ck = json.load(open(os.path.join(GOLDEN, "loot_checkpoint_spec.json")))
    assert [(n, list(s)) for n, s in P.param_spec(7)] == [(n, s) for n, s in ck["params"]]
    assert P.offsets(P.param_spec(7))[-1] == 54712
    import ctypes as C
    buf = (C.c_int64 * 512)()
    n = _lib.load().linr_param_offsets(7, buf, 512)
    assert [int(buf[i]) for i in range(n)] == P.offsets(P.param_spec(7))[:-1]
    flat = P.init_flat(7, seed=1)
    v = P.named_views(flat, 7)
    assert v["scale_mlp.0.0.bias"].abs().max() == 0                      # Linear bias zero (module_utils.py:42-61)
    k = v["upsampler.block_in.0.kernel"]
    assert k.abs().max() <= 1.0 / np.sqrt(27 * 8) + 1e-7                  # ME init bound
    assert v["upsampler.block_in.0.bias"].abs().max() > 0


def test_host_range_coder_is_bitstream_identical_to_oracle():
    from oracle import rc as orc
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 33, 20000):
        cdf = rng.integers(1, 65536, size=n).astype(np.uint16)
        cdf[: n // 10] = 1
        cdf[n // 10: n // 5] = 65535
        sym = (rng.random(n) < (1 - cdf / 65536.0)).astype(np.uint8)
        got = rc.encode_binary(cdf, sym)
        rows = np.stack([np.zeros(n, np.uint16), cdf, np.zeros(n, np.uint16)], axis=1).reshape(-1, 3)
        assert got == orc.encode_u16(rows, sym.astype(np.int16))
        assert (rc.decode_binary(cdf, got, n) == sym).all()
    # shared-row coder (model.bin) incl. the reference's CDF quirk row
    row = model_compression.laplace_cdf_row(129.0, 6.0, 8)
    sym = np.clip(np.rint(rng.laplace(129, 6, size=5000)), 0, 255).astype(np.int16)
    b = rc.encode_shared(row, sym)
    assert b == orc.encode_u16(row, sym)
    assert (rc.decode_shared(row, b, len(sym)) == sym).all()
    # torchac-shaped entry points
    p = rng.random(1000).astype(np.float32)
    cdf_f = np.stack([np.zeros(1000, np.float32), 1 - p, np.ones(1000, np.float32)], 1)
    s16 = (rng.random(1000) < p).astype(np.int16)
    bs = rc.encode_float_cdf(torch.from_numpy(cdf_f), torch.from_numpy(s16))
    assert bs == orc.encode_float_cdf(cdf_f, s16)
    assert (rc.decode_float_cdf(torch.from_numpy(cdf_f), bs).numpy() == s16).all()


def test_model_decompress_matches_reference_flow_fixture():
    g = np.load(os.path.join(GOLDEN, "net_tiny.npz"), allow_pickle=False)
    enc = dict(enc_mode=int(g["q_enc_mode"]), final_bytes=g["q_bytes"].tobytes(), mu=float(g["q_mu"]), b=float(g["q_b"]),
               min_param=float(g["q_min"]), max_param=float(g["q_max"]), bitdepth=8)
    rec = model_compression.decompress_model(enc, len(g["q_recon"]), device="cpu")
    np.testing.assert_array_equal(rec.numpy(), g["q_recon"])
    from oracle import linr_oracle as O
    assert (model_compression.laplace_cdf_row(129.0, 6.0) == O.cdf_float_to_u16(O.laplace_cdf_row_float(129.0, 6.0))).all()


def test_bitstream_containers_round_trip():
    parts = [b"", b"a", os.urandom(300)]
    assert codec.unpack_bitstream(codec.pack_bitstream(parts)) == parts
    from oracle import linr_oracle as O
    assert codec.pack_bitstream(parts) == O.pack_bitstream(parts)
    lows = [np.array([[1, 2, 3], [4, 5, 255]]), np.array([[0, 0, 0]])]
    mins = [np.array([5, -6, 7]), np.array([0, 0, 0])]
    buf = codec.pack_low_xyz(lows, mins)
    assert buf == O.pack_low_xyz(lows, mins)
    l2, m2 = codec.unpack_low_xyz(buf)
    assert (l2[0] == lows[0]).all() and (l2[1] == lows[1]).all() and (m2 == np.array(mins)).all()
    with pytest.raises(AssertionError):
        codec.pack_low_xyz([np.array([[256, 0, 0]])], [np.zeros(3)])      # test_utils.py:221


def test_ply_io_round_trip(tmp_path):
    pts = synth.make_sequence("tiny", 1)[0].numpy()
    p = str(tmp_path / "a.ply")
    pointio.write_ply_ascii(p, pts)
    assert (pointio.read_ply(p) == pts).all()
    rec = np.zeros(len(pts), dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1")])
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    pb = str(tmp_path / "b.ply")
    with open(pb, "wb") as f:
        f.write(f"ply\nformat binary_little_endian 1.0\nelement vertex {len(pts)}\nproperty float x\nproperty float y\n"
                f"property float z\nproperty uchar red\nend_header\n".encode())
        f.write(rec.tobytes())
    assert (pointio.read_ply(pb) == pts).all()
    np.save(str(tmp_path / "c.npy"), pts)
    d = pointio.PointDirectory(str(tmp_path), "ply")
    assert len(d) == 2 and (d[0] == pts).all()


def test_cli_flags_and_gop_partition():
    ref_flags = {"others_epoch": 100, "first_epoch": 100, "gop_size": 4, "frame_num": 4, "learning_rate": 0.01, "gamma": 0.992,
                 "min_lr": 4e-4, "decay_rate": 1e-4, "step_size": 32, "scale_num": None, "min_point_num": 64, "load": "False",
                 "pretrain_path": None, "write_pth": "True", "seed": 8807, "delete_cache": "False", "write_real_bitstream": "False",
                 "check_freq": 5, "ori_dtype": "ply", "handle_dir": "tmp/test_pc", "model_path": None, "result_dir": "output/test_pc",
                 "hidden_channel_mlp": 24, "mlp_out_channel": 10, "hidden_channel_conv": 8, "block_layers": 1, "model_bitdepth": 8,
                 "overfit": "False", "mid_test": "False", "encode": "False", "encode_dir": "result_enc/test_pc", "decode": "True",
                 "decode_dir": "result_dec/test_pc"}          # main.py:480-534
    a = cli.build_parser().parse_args([])
    for k, v in ref_flags.items():
        assert getattr(a, k) == v, k
    assert cli.gop_ranges(96, 32) == [list(range(0, 32)), list(range(32, 64)), list(range(64, 96))]
    assert cli.gop_ranges(5, 4) == [[0, 1, 2, 3], [4]]               # main.py:81-88


def test_model_class_state_dict_contract():
    from linr_pcgc_b200.model import LINR_PCGC_Model
    m = LINR_PCGC_Model({"scale_num": 7, "in_channel": 7, "hidden_channel_conv": 8, "block_layers": 1, "outstage": 8, "instage": 1})
    sd = m.state_dict()
    ck = json.load(open(os.path.join(GOLDEN, "loot_checkpoint_spec.json")))
    assert [(k, list(v.shape)) for k, v in sd.items()] == [(n, s) for n, s in ck["params"]]
    assert sum(p.numel() for p in m.parameters()) == 54712
    other = {k: torch.full_like(v, 0.5) for k, v in sd.items()}
    m.load_state_dict(other)
    assert float(m.flat.min()) == 0.5 and float(m.flat.max()) == 0.5
    with pytest.raises(RuntimeError):
        m.load_state_dict({"scale_emb.weight": torch.zeros(7, 8)})
    with pytest.raises(ValueError):
        LINR_PCGC_Model({"scale_num": 7, "in_channel": 7, "hidden_channel_conv": 16, "block_layers": 1, "outstage": 8, "instage": 1})
    with pytest.raises(_lib.LinrError):    # no CPU path
        m.forward({"coord": torch.zeros(4, 3, dtype=torch.int32), "occ_lst": [torch.zeros(4, 1)] * 8,
                   "offset_tensor": torch.zeros(4, 7), "scale_idx": 0})


def test_sharding_plans():
    assert D.plan_gops(3, 2) == [[1], [2]]
    assert D.plan_gops(3, 8)[:3] == [[1], [2], []]
    assert D.plan_gops(4, 2, first_is_seed=False) == [[0, 2], [1, 3]]
    # stage split: contiguous, disjoint, covering [0,8) for every group size
    for parts in range(1, 9):
        rs = [D.stage_range(parts, p) for p in range(parts)]
        assert rs[0][0] == 0 and rs[-1][1] == 8 and all(a[1] == b[0] for a, b in zip(rs, rs[1:])) and all(lo < hi for lo, hi in rs)
    assert [D.stage_range(8, p) for p in (0, 7)] == [(0, 1), (7, 8)] and D.stage_range(2, 1) == (4, 8)
    with pytest.raises(ValueError):
        D.stage_range(9, 0)
    # the 96-frame / gop-32 job of BASELINE.json configs[2]: GOP 0 on every rank first, then GOPs 1, 2 on two halves
    assert D.plan_job(3, 8) == [[(list(range(8)), [0])], [([0, 1, 2, 3], [1]), ([4, 5, 6, 7], [2])]]
    assert D.plan_job(3, 4) == [[([0, 1, 2, 3], [0])], [([0, 1], [1]), ([2, 3], [2])]]
    assert D.plan_job(3, 2) == [[([0, 1], [0])], [([0], [1]), ([1], [2])]]
    assert D.plan_job(3, 1) == [[([0], [0])], [([0], [1, 2])]]
    assert D.plan_job(2, 8) == [[(list(range(8)), [0])], [(list(range(8)), [1])]]          # Owlii: 64 frames = 2 GOPs
    assert D.plan_job(1, 4) == [[([0, 1, 2, 3], [0])]]
    ph = D.plan_job(6, 3)
    assert sorted(g for _, gops in ph[1] for g in gops) == [1, 2, 3, 4, 5] and [r for r, _ in ph[1]] == [[0], [1], [2]]
    assert sorted(sum([D.frame_share(32, 4, p) for p in range(4)], [])) == list(range(32))


def test_lr_schedule_matches_torch_adam_steplr_across_gops():
    """The host-side schedule against the classes the reference uses (torch.optim.Adam + StepLR stepped per frame,
    main.py:231-252,319-321; min_lr floor per epoch, main.py:433-437; a later GOP loads the optimizer state and builds a
    fresh StepLR, main.py:241-252).  Two StepLR boundaries, the floor, and a GOP boundary whose step count is NOT a
    multiple of step_size are crossed; the lr used at every optimiser step must be equal."""
    from linr_pcgc_b200.trainer import OptimState, sched_after_step, sched_end_epoch, sched_new_gop
    lr0, gamma, step_size, min_lr, frames = 6e-4, 0.7, 5, 4e-4, 4

    def torch_run(epochs_per_gop):
        w = torch.nn.Parameter(torch.ones(3))
        used, sd = [], None
        for ep_n in epochs_per_gop:
            opt = torch.optim.Adam([w], lr=lr0, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
            if sd is not None:
                opt.load_state_dict(sd)
            sch = torch.optim.lr_scheduler.StepLR(opt, step_size=step_size, gamma=gamma)
            for _ in range(ep_n):
                for _ in range(frames):
                    opt.zero_grad()
                    (w * w).sum().backward()
                    used.append(opt.param_groups[0]["lr"])
                    opt.step()
                    sch.step()
                for g in opt.param_groups:
                    if g["lr"] < min_lr:
                        g["lr"] = min_lr
            sd = opt.state_dict()
        return used

    def ours(epochs_per_gop):
        st = OptimState(torch.zeros(1), torch.zeros(1), torch.zeros(1), 0, 0, lr0)
        used = []
        for gi, ep_n in enumerate(epochs_per_gop):
            if gi:
                sched_new_gop(st)
            for _ in range(ep_n):
                for _ in range(frames):
                    used.append(st.lr)
                    st.step += 1
                    sched_after_step(st, step_size, gamma)
                sched_end_epoch(st, min_lr)
        return used

    want, got = torch_run([3, 3]), ours([3, 3])          # 12 steps per GOP: 12 % 5 != 0 -> the decay phase restarts in GOP 1
    assert len(want) == 24 and len(set(want)) >= 4       # decays, the floor and the restart all show up
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    assert min(want) >= min_lr * gamma and any(abs(x - min_lr) < 1e-15 for x in want)


def test_gop_container_round_trips_and_matches_the_directory_layout(tmp_path):
    """One-file container (container.py) <-> the reference's directory layout (encoder.py:84-146): same payload bytes,
    ragged scale counts, both CDF versions of the side info."""
    from linr_pcgc_b200 import container, pipeline
    rng = np.random.default_rng(3)
    fb = [[rng.bytes(int(rng.integers(1, 200))) for _ in range(ns)] for ns in (3, 3, 2, 1)]
    for cdfv in (1, 2):
        side = dict(mu=129.0, b=6.0, min_param=-0.75, max_param=0.5, enc_mode=2, bitdepth=8)
        if cdfv != 1:
            side["cdf_version"] = cdfv
        enc = pipeline.EncodedGop(3, side, rng.bytes(777), 777 * 8 + 82.0, rng.bytes(99), fb, [1000, 1001, 1002, 17])
        data = container.pack(enc)
        back = container.unpack(data)
        assert back == enc and back.side_info.get("cdf_version", 1) == cdfv
        assert container.write(enc, str(tmp_path / f"g{cdfv}.linr")) == len(data) and container.read(str(tmp_path / f"g{cdfv}.linr")) == enc
        # container -> directory -> container: nothing but the index changes
        pipeline.write_gop(back, str(tmp_path / f"dir{cdfv}"))
        again = pipeline.read_gop(str(tmp_path / f"dir{cdfv}"), 3, 4)
        assert again.frame_bytes == fb and again.model_bytes == enc.model_bytes and again.low_enc_bytes == enc.low_enc_bytes
        assert {k: again.side_info[k] for k in side} == side
        overhead = len(data) - sum(len(b) for f in fb for b in f) - 777 - 99
        assert overhead == container._HEAD.size + 4 * 6 + 16 * (2 + 9) == 268      # header + per-frame meta + index
    with pytest.raises(ValueError):
        container.unpack(b"garbage" * 20)
    with pytest.raises(ValueError):
        container.unpack(data[:-5])


def test_checkpoint_is_loadable_by_torch_adam_like_the_reference(tmp_path):
    """main.py:241-246 of the reference: `optimizer.load_state_dict(ckpt['optimizer_state_dict'])` on an Adam over
    model.parameters() -- a model.pth written by save_checkpoint must pass through that, and come back through
    load_checkpoint unchanged (interoperability in both directions)."""
    from linr_pcgc_b200 import main as M, params as P
    from linr_pcgc_b200.trainer import OptimState
    S = 3
    n = P.offsets(P.param_spec(S))[-1]
    g = torch.Generator().manual_seed(1)
    st = OptimState(torch.randn(n, generator=g), torch.randn(n, generator=g), torch.rand(n, generator=g), 37, 5, 7.5e-3)
    path = str(tmp_path / "model.pth")
    M.save_checkpoint(path, st, S, 9, 0.25, 8, 0.01, 1e-4)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    spec = P.param_spec(S)
    assert list(ck["model"].keys()) == [k for k, _ in spec] and all(tuple(ck["model"][k].shape) == tuple(shp) for k, shp in spec)
    plist = [torch.nn.Parameter(ck["model"][k].clone()) for k, _ in spec]          # what model.parameters() yields, in order
    opt = torch.optim.Adam(plist, lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    opt.load_state_dict(ck["optimizer_state_dict"])
    assert opt.param_groups[0]["lr"] == 7.5e-3 and opt.param_groups[0]["initial_lr"] == 0.01
    m_back = torch.cat([opt.state[p]["exp_avg"].reshape(-1) for p in plist])
    assert torch.equal(m_back, st.m) and float(opt.state[plist[0]]["step"]) == 37.0
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=32, gamma=0.992)          # main.py:252 on the loaded optimizer
    assert sched.get_last_lr() == [7.5e-3]
    back, S2 = M.load_checkpoint(path, "cpu")
    assert S2 == S and torch.equal(back.params, st.params) and torch.equal(back.m, st.m) and torch.equal(back.v, st.v)
    assert back.step == 37 and back.lr == 7.5e-3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from linr_pcgc_b200.trainer import OptimState
    n = 1000
    if rank == 0:
        st = OptimState(torch.arange(n, dtype=torch.float32), torch.ones(n), torch.full((n,), 2.0), 17, 19, 1.25e-3)
    else:
        st = OptimState(torch.empty(n), torch.empty(n), torch.empty(n), 0, 0, 0.0)
    st = D.broadcast_state(st, 0)
    ok = bool((st.params == torch.arange(n, dtype=torch.float32)).all()) and st.step == 17 and st.sched_step == 19 and st.lr == 1.25e-3
    # sub-groups of a plan_job schedule: created collectively, a singleton group needs no communicator
    phases = D.plan_job(3, world)
    D.make_groups(phases)
    ok = ok and D.group_for([0, 1]) is None and D.group_for([rank]) is not None
    g = torch.full((n,), float(rank + 1))
    dist.all_reduce(g, group=D.group_for(phases[0][0][0]))   # the stage-split gradient sum over GOP 0's ranks
    ok = ok and bool((g == 3.0).all())
    g1 = torch.full((4,), float(rank + 1))
    dist.all_reduce(g1, group=D.group_for([rank]))           # phase 1: one rank per GOP, nothing crosses ranks
    ok = ok and bool((g1 == float(rank + 1)).all())
    gathered = D.gather_bytes([bytes([rank])] * 2)
    if rank == 0:
        ok = ok and gathered == [[b"\x00", b"\x00"], [b"\x01", b"\x01"]]
    out[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_state_broadcast_and_grad_allreduce():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_me_shim_installs_and_refuses_cpu_tensors():
    """linr_pcgc_b200.shim registers `MinkowskiEngine` / `torchac`; modules build on the host with the reference's
    parameter shapes, but there is no CPU compute path."""
    import importlib
    import linr_pcgc_b200.shim as shim
    saved = {k: sys.modules.get(k) for k in ("MinkowskiEngine", "torchac")}
    try:
        shim.install(force=True)
        ME = importlib.import_module("MinkowskiEngine")
        ta = importlib.import_module("torchac")
        assert callable(ta.encode_float_cdf) and callable(ta.decode_float_cdf)
        for name in ("SparseTensor", "MinkowskiConvolution", "MinkowskiReLU", "MinkowskiPruning", "cat", "utils"):
            assert hasattr(ME, name), name
        c3 = ME.MinkowskiConvolution(5, 8, kernel_size=3, stride=1, bias=True, dimension=3)
        c1 = ME.MinkowskiConvolution(8, 4, kernel_size=1, stride=1, bias=True, dimension=3)
        assert tuple(c3.kernel.shape) == (27, 5, 8) and tuple(c3.bias.shape) == (1, 8) and tuple(c1.kernel.shape) == (8, 4)
        assert sorted(k for k, _ in c3.named_parameters()) == ["bias", "kernel"]
        assert float(c3.kernel.abs().max()) <= 1.0 / np.sqrt(5 * 27) + 1e-7
        xyz = torch.tensor([[0, 0, 0], [0, 0, 1], [1, 0, 0]], dtype=torch.int32)
        C, F = ME.utils.sparse_collate([xyz], [torch.ones(3, 5)])
        assert C.shape == (3, 4) and int(C[:, 0].abs().sum()) == 0 and torch.equal(C[:, 1:], xyz)
        with pytest.raises(RuntimeError):
            ME.SparseTensor(features=F, coordinates=C)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
