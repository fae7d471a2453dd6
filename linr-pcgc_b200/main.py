"""Command line of the codec: the reference's `main.py` flags (main.py:480-534) driving the B200 path.

    python -m linr_pcgc_b200.main --overfit True --encode True --decode True --ori_dir <plys> --frame_num 96 \
        --gop_size 32 --first_epoch 10 --others_epoch 10 --result_dir out --encode_dir enc --decode_dir dec
    torchrun --nproc-per-node 8 -m linr_pcgc_b200.main ...        # GOP 0 stage-split over all GPUs, then the other GOPs on
                                                                   # groups of GPUs (linr_pcgc_b200.dist.plan_job)

Every flag of the reference is accepted with the same name, type and default (booleans are the strings
'True'/'False' as there); `--synthetic <shape>` replaces `--ori_dir` by the seeded generator of synth.py.  Outputs
keep the reference's layout: <result_dir>/<gop>/model.pth + result.json, <encode_dir>/<gop>/{side_info.json,
bins/model.bin, bins/low_enc_bytes.bin, bins/frame%04d_scale%d.bin}, <decode_dir>/<gop>/*.ply
(main.py:69-120, encoder.py:57-156, decoder.py:51-147).
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import time
from typing import List, Optional

import numpy as np
import torch

from . import dist as D
from . import params as P
from . import pipeline, pointio, synth
from .trainer import GopTrainer, OptimState

logger = logging.getLogger("LINR_PCGC")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser("LINR-PCGC")
    p.add_argument("--others_epoch", default=100, type=int)
    p.add_argument("--first_epoch", default=100, type=int)
    p.add_argument("--gop_size", type=int, default=4)
    p.add_argument("--frame_num", type=int, default=4)
    p.add_argument("--learning_rate", default=0.01, type=float)
    p.add_argument("--gamma", type=float, default=0.992)
    p.add_argument("--min_lr", type=float, default=4e-4)
    p.add_argument("--decay_rate", type=float, default=1e-4)
    p.add_argument("--step_size", type=int, default=32)
    p.add_argument("--scale_num", type=int)
    p.add_argument("--min_point_num", type=int, default=64)
    p.add_argument("--load", default="False", type=str)
    p.add_argument("--pretrain_path", type=str)
    p.add_argument("--write_pth", type=str, default="True")
    p.add_argument("--seed", type=int, default=8807)
    p.add_argument("--delete_cache", type=str, default="False")
    p.add_argument("--write_real_bitstream", type=str, default="False")
    p.add_argument("--check_freq", type=int, default=5)
    p.add_argument("--ori_dir", type=str, default=None)
    p.add_argument("--ori_dtype", type=str, default="ply")
    p.add_argument("--handle_dir", type=str, default="tmp/test_pc")
    p.add_argument("--model_path", type=str, default=None)
    p.add_argument("--result_dir", type=str, default="output/test_pc")
    p.add_argument("--hidden_channel_mlp", type=int, default=24)
    p.add_argument("--mlp_out_channel", type=int, default=10)
    p.add_argument("--hidden_channel_conv", type=int, default=8)
    p.add_argument("--block_layers", type=int, default=1)
    p.add_argument("--model_bitdepth", type=int, default=8)
    p.add_argument("--overfit", type=str, default="False")
    p.add_argument("--mid_test", type=str, default="False")
    p.add_argument("--encode", type=str, default="False")
    p.add_argument("--encode_dir", type=str, default="result_enc/test_pc")
    p.add_argument("--decode", type=str, default="True")
    p.add_argument("--decode_dir", type=str, default="result_dec/test_pc")
    # additions of this implementation
    p.add_argument("--synthetic", type=str, default=None, help="generate the sequence with synth.py (loot, owlii, mvub9, ...)")
    p.add_argument("--apply_seed", type=str, default="False", help="seed the parameter init with --seed (the reference never seeds)")
    return p


class Sequence:
    """Frame source: a directory of PLY/NPY files or the synthetic generator."""

    def __init__(self, args, device):
        self.device = device
        if args.synthetic:
            self.src = None
            self.shape = args.synthetic
        else:
            if not args.ori_dir:
                raise SystemExit("--ori_dir (or --synthetic) is required")
            self.src = pointio.PointDirectory(args.ori_dir, args.ori_dtype)

    def points(self, idx: List[int]) -> List[torch.Tensor]:
        if self.src is None:
            return [synth.make_sequence(self.shape, 1, device=self.device, start=i)[0] for i in idx]
        return [torch.from_numpy(self.src[i]).pin_memory().to(self.device, non_blocking=True) for i in idx]


def gop_ranges(frame_num: int, gop_size: int) -> List[List[int]]:
    """main.py:81-88."""
    return [list(range(i, min(i + gop_size, frame_num))) for i in range(0, frame_num, gop_size)]


def save_checkpoint(path: str, state: OptimState, scale_num: int, epoch: int, loss: float, bitdepth: int,
                    learning_rate: float = 0.01, weight_decay: float = 1e-4):
    """model.pth in the reference's layout (main.py:365-374): 'model' = state_dict by the reference's tensor names,
    'optimizer_state_dict' = what torch.optim.Adam.state_dict() gives for model.parameters() in order (per-tensor step /
    exp_avg / exp_avg_sq, one param group with lr and initial_lr), so the reference's own
    `optimizer.load_state_dict(ckpt['optimizer_state_dict'])` (main.py:243-246) accepts a checkpoint written here."""
    cpu = lambda t: t.detach().cpu()
    views = P.named_views(cpu(state.params), scale_num)
    m_views, v_views = P.named_views(cpu(state.m), scale_num), P.named_views(cpu(state.v), scale_num)
    names = [n for n, _ in P.param_spec(scale_num)]
    opt_state = {i: {"step": torch.tensor(float(state.step)), "exp_avg": m_views[n].clone(), "exp_avg_sq": v_views[n].clone()}
                 for i, n in enumerate(names)}
    group = {"lr": state.lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": weight_decay, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
             "initial_lr": learning_rate, "params": list(range(len(names)))}
    torch.save({"model": {n: views[n].clone() for n in names}, "epoch": epoch, "loss": loss, "bitdepth": bitdepth, "scale_num": scale_num,
                "optimizer_state_dict": {"state": opt_state, "param_groups": [group]},
                "linr_b200": {"sched_step": state.sched_step}}, path)


def load_checkpoint(path: str, device) -> (OptimState, int):
    ck = torch.load(path, map_location="cpu", weights_only=False)
    model = ck["model"]
    scale_num = int(ck.get("scale_num", model["scale_emb.weight"].shape[0]))
    flat = torch.cat([model[n].reshape(-1).float() for n, _ in P.param_spec(scale_num)]).to(device)
    opt = ck.get("optimizer_state_dict", {})
    if opt.get("format") == "linr_b200_flat_adam":   # round-1 checkpoints of this implementation
        st = OptimState(flat, opt["m"].to(device), opt["v"].to(device), int(opt["step"]), int(opt["sched_step"]), float(opt["lr"]))
    elif "state" in opt:  # a checkpoint written by the reference: torch.optim.Adam state per tensor, parameters() order
        ms = torch.cat([opt["state"][i]["exp_avg"].reshape(-1).float() for i in range(len(opt["state"]))]).to(device)
        vs = torch.cat([opt["state"][i]["exp_avg_sq"].reshape(-1).float() for i in range(len(opt["state"]))]).to(device)
        step = int(float(opt["state"][0]["step"]))
        st = OptimState(flat, ms, vs, step, step, float(opt["param_groups"][0]["lr"]))
    else:
        st = OptimState(flat, torch.zeros_like(flat), torch.zeros_like(flat), 0, 0, 0.01)
    return st, scale_num


def overfit_one_gop(args, seq: Sequence, group: List[int], epochs: int, seed_state: Optional[OptimState], device,
                    ranks: Optional[List[int]] = None):
    """main.py:122-455 for one GOP; returns (OptimState, scale_num).  `ranks`: the ranks that share this GOP -- each
    computes its stages of every frame (dist.stage_range), one gradient all-reduce per frame; the first one writes."""
    name = f"gop_{group[0]}_{group[-1]}"
    gdir = os.path.join(args.result_dir, name)
    os.makedirs(gdir, exist_ok=True)
    ranks = ranks if ranks and len(ranks) > 1 else None
    parts, part = (len(ranks), ranks.index(D.rank())) if ranks else (1, 0)
    pg = D.group_for(ranks) if ranks else None
    writer = ranks is None or part == 0
    frames = pipeline.prepare_gop(seq.points(group), args.scale_num, args.min_point_num, device)
    S = args.scale_num or frames[0].n_scales
    args.scale_num = S
    seed = args.seed if args.apply_seed == "True" else None
    if seed_state is None and ranks and seed is None:
        # the reference never seeds (main.py:504 is parsed and unused); the members of a stage split must still start
        # from ONE random model: the first rank draws the seed
        t = torch.randint(0, 2 ** 31 - 1, (1,), device=device)
        torch.distributed.broadcast(t, ranks[0], group=pg)
        seed = int(t.item())
    tr = GopTrainer(S, device, args.learning_rate, args.gamma, args.step_size, args.min_lr, args.decay_rate, seed=seed,
                    state=seed_state.clone() if seed_state is not None else None,
                    max_rows=max(f.tables.n_rows for f in frames), ranks=ranks, group=pg)
    results, t_train = [], 0.0
    for ep in range(epochs):
        torch.cuda.synchronize()
        t0 = time.time()
        loss = tr.fit(frames, 1)[0]
        torch.cuda.synchronize()
        t_train += time.time() - t0
        rec = {"epoch": ep, "loss": loss, "train_time": t_train, "train_time_avg": t_train / len(group)}
        if args.mid_test == "True" and writer and (ep < 10 or ep % args.check_freq == 0):
            enc = pipeline.encode_gop(frames, tr.state.params, S, args.model_bitdepth)
            rec.update({"real_bpp_all": enc.bpp, "model_bpp": enc.model_bits / sum(enc.point_nums),
                        "xyzlow_bpp": 8 * len(enc.low_enc_bytes) / sum(enc.point_nums), "enc_mode": enc.side_info["enc_mode"]})
        results.append(rec)
        logger.info(f"{name} epoch {ep} loss {loss:.5f} train_time {t_train:.2f}s lr {tr.state.lr:.3e}")
        if writer:
            with open(os.path.join(gdir, "result.json"), "w") as f:
                json.dump(results, f, indent=4)
    if args.write_pth == "True" and writer:
        save_checkpoint(os.path.join(gdir, "model.pth"), tr.state, S, epochs - 1, results[-1]["loss"] if results else 0.0,
                        args.model_bitdepth, args.learning_rate, args.decay_rate)
    return tr.state, S


def run(args) -> None:
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=device)
    os.makedirs(args.result_dir, exist_ok=True)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(name)s - %(levelname)s - %(message)s",
                        handlers=[logging.StreamHandler(), logging.FileHandler(os.path.join(args.result_dir, f"info_rank{D.rank()}.log"))])
    seq = Sequence(args, device)
    groups = gop_ranges(args.frame_num, args.gop_size)
    names = [f"gop_{g[0]}_{g[-1]}" for g in groups]
    phases = D.plan_job(len(groups), D.world())
    D.make_groups(phases)
    # GOPs whose files this rank writes: the first rank of the group that trained them
    mine = [g for phase in phases for ranks, gops in phase for g in gops if ranks[0] == D.rank()]

    if args.overfit == "True":
        seed_state = None
        if args.pretrain_path and os.path.exists(str(args.pretrain_path)):
            seed_state, args.scale_num = load_checkpoint(args.pretrain_path, device)
        first = phases[0][0][0]
        state0 = None
        if D.rank() in first:
            state0, S = overfit_one_gop(args, seq, groups[0], args.first_epoch, seed_state, device, ranks=first)
        if D.world() > len(first):   # more ranks than one GOP can use: ship GOP 0's state to the others
            sn = torch.tensor([args.scale_num or 0], device=device)
            torch.distributed.broadcast(sn, 0)
            args.scale_num = int(sn.item())
            if state0 is None:
                n = P.offsets(P.param_spec(args.scale_num))[-1]
                state0 = OptimState(torch.empty(n, device=device), torch.empty(n, device=device), torch.empty(n, device=device), 0, 0, 0.0)
            state0 = D.broadcast_state(state0, 0)
        for ranks, gops in (phases[1] if len(phases) > 1 else []):
            if D.rank() in ranks:
                for g in gops:
                    overfit_one_gop(args, seq, groups[g], args.others_epoch, state0, device, ranks=ranks)
        if D.world() > 1:
            torch.distributed.barrier()

    todo = mine
    if args.encode == "True":
        for g in todo:
            st, S = load_checkpoint(os.path.join(args.result_dir, names[g], "model.pth"), device)
            frames = pipeline.prepare_gop(seq.points(groups[g]), S, args.min_point_num, device)
            enc = pipeline.encode_gop(frames, st.params, S, args.model_bitdepth)
            pipeline.write_gop(enc, os.path.join(args.encode_dir, names[g]))
            logger.info(f"{names[g]} encoded: {enc.bpp:.5f} bpp over {sum(enc.point_nums)} points")
    if args.decode == "True" and os.path.isdir(args.encode_dir):
        for g in todo:
            gdir = os.path.join(args.encode_dir, names[g])
            if not os.path.isdir(gdir):
                continue
            S = args.scale_num or load_checkpoint(os.path.join(args.result_dir, names[g], "model.pth"), device)[1]
            enc = pipeline.read_gop(gdir, S, len(groups[g]))
            dec = pipeline.decode_gop(enc, device)
            os.makedirs(os.path.join(args.decode_dir, names[g]), exist_ok=True)
            ori = seq.points(groups[g])
            for i, (d, o) in enumerate(zip(dec, ori)):
                ref = torch.unique(o.to(torch.int64), dim=0).to(torch.int32)   # sorted original (decoder.py:135-139)
                assert d.shape == ref.shape and bool((d == ref).all()), f"{names[g]} frame {i}: decode is not lossless"
                pointio.write_ply_ascii(os.path.join(args.decode_dir, names[g], f"frame{i:04d}_dec.ply"), d.cpu().numpy())
            logger.info(f"{names[g]} decoded losslessly ({len(dec)} frames)")
    if D.world() > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main(argv=None):
    args = build_parser().parse_args(argv)
    print(args)
    run(args)


if __name__ == "__main__":
    main()
