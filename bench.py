#!/usr/bin/env python
"""bench.py — overfit-plus-encode seconds per frame of the LINR-PCGC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A STEP is the whole north-star job (BASELINE.json configs[1] / configs[2]): a loot-shaped sequence of `--gops` (3) GOPs of
`--frames` (32) synthetic 10-bit frames of ~780k points = 96 frames.  GOP 0 is overfitted first (`--epochs` (10) passes of
per-frame forward + backward + Adam) because its parameters, Adam moments and learning rate seed the later GOPs
(main.py:102-104); then the other GOPs; every GOP is followed by 8-bit model quantisation and the real encode of its frames
(network forward, 16-bit CDFs to the host, range coder).
  N = 1 : the three GOPs one after the other on one GPU.
  N > 1 : dist.plan_job -- GOP 0 stage-split over all N GPUs (each rank computes its share of the 8 autoregressive stages of
          every frame; the first rank also owns SCE + block_in, broadcasts its output and receives the reduced gradient
          of it; ONE all-reduce of the flat 219 kB gradient per frame, identical fused Adam on every rank: the reference's
          one-optimiser-step-per-frame semantics are kept), then GOPs 1 and 2 at the same time on two halves of the GPUs,
          each stage-split over its half.  Fixed total work: "scaling": "strong".
  value : inputs (prepared frames) resident in HBM when the timed region starts; whole-job seconds / 96 frames.
  e2e   : the same job fed pinned HOST point arrays: H2D copy, octree / kernel-map preparation, overfit, encode, bitstreams
          collected on rank 0 -- all inside the timed region.
  replica: (extra key) the round-1 measure, one independent GOP per GPU with no collective (weak scaling).
`--mode gop` times only that replica measure; `--decode-only` times BASELINE.json configs[4] (decode s/frame).
`--impl reference` times the CPU restatement of the reference (oracle/, torch CPU, all host threads) on one whole
frame-iteration + one frame encode of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# before the CUDA context exists (linr-pcgc_b200/__init__.py says why): one hardware queue per stream
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import torch

METRIC = "overfit_plus_encode_s_per_frame"
UNIT = "s/frame"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default="loot")
    ap.add_argument("--frames", type=int, default=32, help="GOP size (main.py --gop_size)")
    ap.add_argument("--gops", type=int, default=None, help="GOPs of the sequence (default 3 = 96 frames; owlii: 2 = 64 frames)")
    ap.add_argument("--epochs", type=int, default=10, help="first_epoch / others_epoch of the north-star config")
    ap.add_argument("--mode", default="job", choices=["job", "gop"], help="job: the whole sequence; gop: one GOP per GPU (round-1 measure)")
    ap.add_argument("--decode-only", action="store_true", help="BASELINE.json configs[4]: decode s/frame of an encoded GOP")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the replica measure and the C4 / C5 side configs")
    ap.add_argument("--cpu-sample-rows", type=int, default=75000)
    ap.add_argument("--cpu-full-frame", action="store_true", help="cpu baseline on every scale of the frame (default for --impl reference)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.stop_flag, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag.set()
        if self.th:
            self.th.join(timeout=6)
        sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_sample(shape: str, sample_rows: int, threads: int):
    """One frame of the workload for the CPU restatement.  sample_rows > 0: only the coarsest scales whose parent-voxel
    count stays below `sample_rows` (a bounded sample; cost is linear in voxel-passes, the ratio is reported);
    sample_rows <= 0: every scale."""
    from linr_pcgc_b200 import params as P, synth
    from oracle import linr_oracle as O
    torch.set_num_threads(threads)
    pts = synth.make_sequence(shape, 1)[0].numpy()
    fr = O.prepare_frame(pts, None, 64)
    rows = [len(s["coord"]) for s in fr["scales"]]
    total = sum(rows)
    keep, acc = [], 0
    for i in range(len(rows) - 1, -1, -1):
        if sample_rows > 0 and acc + rows[i] > sample_rows and keep:
            break
        keep.append(i)
        acc += rows[i]
    keep.sort()
    sub = dict(fr)
    sub["scales"] = [fr["scales"][i] for i in keep]
    nbrs = [torch.from_numpy(O.nbr27(s["coord"]).astype(np.int64)) for s in sub["scales"]]
    S = len(fr["scales"])
    flat = P.init_flat(S, seed=1)
    return O, sub, nbrs, S, flat, acc, total, fr["point_num"]


def oracle_iteration(O, sub, nbrs, S, flat, m, v, step):
    """One frame-iteration of the reference algorithm on the CPU: forward, backward, Adam (main.py:305-321)."""
    p = flat.clone().requires_grad_(True)
    sd = O.unflatten_params(p, S)
    bits = O.frame_bits(sd, sub, nbrs)
    (bits / sub["point_num"]).backward()
    with torch.no_grad():
        O.adam_step_reference([flat], [p.grad], [m], [v], step=step, lr=0.01)
    return float(bits.detach())


def oracle_encode(O, sub, nbrs, S, flat):
    """One frame encode on the CPU: forward without grad + 8 range-coder streams per scale (encoder.py:158-203)."""
    from oracle import rc
    sd = O.unflatten_params(flat, S)
    nbytes = 0
    with torch.no_grad():
        for sc, nb in zip(sub["scales"], nbrs):
            _, p = O.scale_forward(sd, sc, nb)
            for k in range(8):
                pk = p[:, k].numpy()
                n = len(pk)
                cdf = np.stack([np.zeros(n, np.float32), np.float32(1.0) - pk, np.ones(n, np.float32)], axis=1)
                nbytes += len(rc.encode_float_cdf(cdf, sc["occ"][:, k].astype(np.int16)))
    return nbytes


def cpu_reference_time(args, steps: int, warmup: int, full: bool):
    """s/frame of the CPU restatement = epochs x (one frame-iteration) + (one frame encode).  `full`: every scale of
    the frame is timed (about 40 s per iteration on 16 cores); else a bounded sample of the coarsest scales, scaled
    linearly in voxel-passes.  Both figures are kept apart in the returned dict."""
    threads = os.cpu_count() or 1
    O, sub, nbrs, S, flat, rows, total, point_num = oracle_sample(args.shape, 0 if full else args.cpu_sample_rows, threads)
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    scale = total / rows
    for i in range(warmup):
        oracle_iteration(O, sub, nbrs, S, flat, m, v, i + 1)
    its = []
    for i in range(steps):
        t0 = time.perf_counter()
        oracle_iteration(O, sub, nbrs, S, flat, m, v, warmup + i + 1)
        its.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    oracle_encode(O, sub, nbrs, S, flat)
    t_enc_m = time.perf_counter() - t0
    t_iter_m = float(np.mean(its))
    t_iter, t_enc = t_iter_m * scale, t_enc_m * scale
    s_per_frame = args.epochs * t_iter + t_enc
    what = "every scale" if full else "coarsest scales"
    sample = (f"oracle port (torch CPU), {rows} of {total} voxel-passes of one {args.shape} frame ({what}), "
              f"{steps} frame-iteration(s) + 1 encode timed: {t_iter_m:.2f} s/iteration, {t_enc_m:.2f} s/encode measured"
              + ("" if full else f", scaled by {scale:.2f} to the full frame") + f"; s/frame = {args.epochs} x iteration + encode")
    return {"value": s_per_frame, "iter_s": t_iter, "encode_s": t_enc, "iter_s_measured": t_iter_m, "encode_s_measured": t_enc_m,
            "voxel_passes_timed": rows, "voxel_passes_frame": total, "threads": threads, "sample": sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one whole frame-iteration is ~40 s of CPU work: one timed iteration per step, at most two steps, no warm-up pass
    steps = max(1, min(args.steps, 2))
    t_all0 = time.perf_counter()
    r = cpu_reference_time(args, steps, 0, full=True)
    G = args.gops or (2 if args.shape == "owlii" else 3)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["value"] * args.frames * G * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, 1, G),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "iter_s": r["iter_s"], "encode_s": r["encode_s"], "wall_s": time.perf_counter() - t_all0,
            "note": "CPU restatement of the reference (the reference's own natives, MinkowskiEngine / torchac, are not installable "
                    "offline); a whole frame-iteration is timed, s/frame is iteration x epochs + encode -- context, not a speed-up "
                    "over the reference's GPU path"}
    print(json.dumps(line), flush=True)


def workload_config(args, n, G):
    from linr_pcgc_b200 import dist as D, synth
    bits, pts = synth.SHAPES[args.shape]
    cfg = {"loot": "BASELINE.json configs[1] (N=1) / configs[2] (N>1)", "owlii": "BASELINE.json configs[3]",
           "plumbing": "BASELINE.json configs[0]"}.get(args.shape, "parity-test shape")
    if args.mode == "gop":
        par = "gop%d (one independent GOP per GPU, no collective)" % n
    elif n == 1:
        par = "1 GPU, GOPs one after the other"
    else:
        ph = D.plan_job(G, n)
        par = ("GOP 0 stage-split over %d GPUs (g broadcast, dg reduce, gradient all-reduce per frame), then %s" %
               (len(ph[0][0][0]), "; ".join("GOPs %s stage-split over GPUs %d-%d" % (g, r[0], r[-1]) for r, g in ph[1]) if len(ph) > 1 else "-"))
    return {"workload": f"{args.shape}-shaped synthetic {bits}-bit surface, ~{pts // 1000}k pts/frame, gop_size {args.frames}, "
                        f"{G} GOPs = {G * args.frames} frames, {args.epochs} epochs/GOP, overfit + model quantisation + encode ({cfg})",
            "gop_size": args.frames, "gops": G, "epochs": args.epochs,
            "frames_per_step": args.frames * (n if args.mode == "gop" else G),
            "parallelism": par,
            "l2_policy": "inputs larger than L2: one GOP's resident tables + activations >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------ B200 arm
class Ctx:
    """Everything the measures share: process group, devices, library handles."""

    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        import torch.distributed as dist
        self.dist = dist
        from linr_pcgc_b200 import dist as D
        self.D = D
        self.cores = D.bind_rank_cores()
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from linr_pcgc_b200 import _lib
        self.lib = _lib.load()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, prof_mask=0, serial=False):
        """K calls of fn bracketed by barrier + synchronize on both sides; device time (CUDA events), max over ranks.
        serial: keep every launch on the caller's stream (linr_side_stream_enable(0)) so that the per-class CUDA-event
        times of a fully profiled step do not overlap; the timed region of the headline never uses it."""
        self.lib.linr_prof_enable(prof_mask)
        prev = self.lib.linr_side_stream_enable(0) if serial else None
        try:
            return self._timed(fn, steps)
        finally:
            if prev is not None:
                self.lib.linr_side_stream_enable(prev)

    def _timed(self, fn, steps):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item()), wall, out

    def prof_table(self):
        import ctypes as C
        tab = []
        for c in range(self.lib.linr_prof_classes()):
            ms, n, u = C.c_double(), C.c_int64(), C.c_int64()
            self.lib.linr_prof_read(c, C.byref(ms), C.byref(n), C.byref(u))
            tab.append({"kernel": self.lib.linr_prof_name(c).decode(), "ms": ms.value, "launches": n.value, "units": u.value})
        return tab

    def all_mask(self):
        return (1 << self.lib.linr_prof_classes()) - 1


def mean_occupied(frames):
    pop = 0.0
    for f in frames[:4]:
        m = f.tables.mask.to(torch.int64) & 0x7FFFFFF
        cnt = torch.zeros_like(m)
        for b in range(27):
            cnt += (m >> b) & 1
        pop += float(cnt.double().mean().item())
    return pop / max(1, len(frames[:4]))


def roofline_step(live, pbar, ms):
    """Whole timed region: algorithmic bytes of every kernel class launched in it (launch counters are always on) over
    its duration -- the figure that does not depend on which kernels share the SMs at a given moment."""
    r = roofline_of({"kernel": "conv27<8,8>", "units": 0, "ms": 1, "launches": 0}, pbar, 1.0)
    total = sum(algorithmic_bytes(t["kernel"], pbar) * t["units"] for t in live)
    ach = total / max(ms, 1e-9) / 1e6
    return {"bound": "hbm", "what": "all kernel classes of the timed region (overfit + encode forward passes)", "achieved": ach,
            "unit": "GB/s", "peak": r["peak"], "peak_source": r["peak_source"], "frac": ach / r["peak"],
            "algorithmic_bytes": total}


def add_alone(roof, alone_row, pbar):
    """The same class timed with every launch on one stream (the fully profiled warm-up step): the kernel by itself."""
    if not roof or not alone_row or not alone_row.get("launches"):
        return roof
    a = (algorithmic_bytes(alone_row["kernel"], pbar) * alone_row["units"]) / max(alone_row["ms"], 1e-9) / 1e6
    roof["alone"] = {"achieved": a, "frac": a / roof["peak"], "avg_launch_us": 1e3 * alone_row["ms"] / alone_row["launches"],
                     "note": "one stream (linr_side_stream_enable(0)), warm-up step; `achieved` / `frac` above are live in the timed "
                             "region, where the weight-gradient launches of the second stream share the SMs with this class"}
    return roof


def roofline_of(d, pbar, step_ms):
    """Roofline entry of one kernel class from its live CUDA-event time (measured in the timed region)."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bpu = algorithmic_bytes(d["kernel"], pbar)
    achieved = (bpu * d["units"]) / max(d["ms"], 1e-9) / 1e6 if d["launches"] else 0.0  # GB/s
    return {"bound": "hbm", "kernel": d["kernel"], "achieved": achieved, "peak": peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
            "unit": "GB/s", "frac": achieved / peak,
            "traffic": (NCU_DRAM_BYTES_PER_UNIT[d["kernel"]] * d["units"] / max(1, d["launches"])
                        if d["kernel"] in NCU_DRAM_BYTES_PER_UNIT else None),
            "traffic_source": NCU_TRAFFIC_SOURCE,
            "launches": d["launches"], "avg_launch_us": 1e3 * d["ms"] / max(1, d["launches"]),
            "algorithmic_bytes_per_launch": bpu * d["units"] / max(1, d["launches"]),
            "mean_occupied_neighbours": pbar, "share_of_step": d["ms"] / max(step_ms, 1e-9)}


class Job:
    """The whole sequence on this rank's share of the schedule (dist.plan_job)."""

    def __init__(self, cx: Ctx, shape: str, F: int, G: int, E: int):
        from linr_pcgc_b200 import pipeline, synth
        from linr_pcgc_b200.net import NetRunner
        from linr_pcgc_b200.trainer import GopTrainer
        self.cx, self.shape, self.F, self.G, self.E = cx, shape, F, G, E
        D, dev, rank = cx.D, cx.dev, cx.rank
        self.phases = D.plan_job(G, cx.world)
        D.make_groups(self.phases)
        self.mine = []      # (gop, ranks) in execution order
        for phase in self.phases:
            for ranks, gops in phase:
                if rank in ranks:
                    self.mine += [(g, ranks) for g in gops]
        self.pts_dev = {g: synth.make_sequence(shape, F, device=dev, start=g * F) for g, _ in self.mine}
        self.pts_host = {g: [p.cpu().pin_memory() for p in v] for g, v in self.pts_dev.items()}
        g0 = self.mine[0][0]
        first = pipeline.prepare_gop(self.pts_dev[g0][:1], None, 64, dev)[0]
        self.S = first.n_scales
        if cx.world > 1:     # every rank must agree on the scale count (discovered from frame 0 of GOP 0, main.py:77-78)
            t = torch.tensor([self.S if g0 == 0 else 0], device=dev)
            cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
            self.S = int(t.item())
        self.frames = {g: pipeline.prepare_gop(v, self.S, 64, dev) for g, v in self.pts_dev.items()}
        self.max_rows = max(f.tables.n_rows for v in self.frames.values() for f in v)
        self.runner = NetRunner(self.S, self.max_rows, dev, train=False)
        self.trainers = {}
        for g, ranks in self.mine:
            key = tuple(ranks)
            if key not in self.trainers:
                self.trainers[key] = GopTrainer(self.S, dev, seed=8807, max_rows=self.max_rows, ranks=ranks, group=D.group_for(ranks))
        self.pipeline = pipeline

    def step(self, from_host: bool = False, encode: bool = True):
        """One whole job; returns {gop: EncodedGop} for the GOPs whose group this rank leads."""
        cx, D = self.cx, self.cx.D
        out, state0 = {}, None
        bg_prep = from_host and os.environ.get("LINR_BENCH_BG_PREP", "1") != "0"
        if bg_prep and not hasattr(self, "preparer"):
            self.preparer = self.pipeline.GopPreparer(cx.dev)
        for i, (g, ranks) in enumerate(self.mine):
            if not from_host:
                frames = self.frames[g]
            elif not bg_prep:
                frames = self.pipeline.prepare_gop(self.pts_host[g], self.S, 64, cx.dev)
            else:
                # end to end: the rank's first GOP is uploaded and prepared in line, every later one in the background
                # (pipeline.GopPreparer) while the GOP before it is overfitted
                frames = self.preparer.collect() if i > 0 else self.pipeline.prepare_gop(self.pts_host[g], self.S, 64, cx.dev)
                if i + 1 < len(self.mine):
                    self.preparer.submit(self.pts_host[self.mine[i + 1][0]], self.S, 64)
            tr = self.trainers[tuple(ranks)]
            if g == 0:
                tr.reset(seed=8807)
            else:
                tr.reset(state=state0.clone())          # later GOPs start from GOP 0's state (main.py:102-104,241-246)
            tr.fit(frames, self.E)
            if g == 0:
                state0 = tr.state.clone() if len(self.mine) > 1 else tr.state
            if encode:
                enc = self.pipeline.encode_gop_shared(frames, tr.state.params, self.S, 8, self.runner, ranks, D.group_for(ranks))
                if enc is not None:
                    out[g] = enc
        return out

    def h2d_bytes(self):
        return sum(int(p.numel()) * 4 for v in self.pts_host.values() for p in v)

    def d2h_bytes(self):
        rows = [f.tables.n_rows for v in self.frames.values() for f in v]
        return sum(r * (8 * 2 + 1) for r in rows) + 8 * len(rows) * self.E + 54712 * len(self.frames)


def decode_check(cx, pipeline, enc, frames, pts_dev, S, nd=16):
    """Decode the first frames of a GOP outside the headline region: lossless check (decoder.py:140) + throughput."""
    nd = min(len(frames), nd)
    sub = pipeline.EncodedGop(enc.scale_num, enc.side_info, enc.model_bytes, enc.model_bits,
                              pipeline.codec.pack_low_xyz([f.scale_coords(S - 1).cpu().numpy() for f in frames[:nd]],
                                                          [f.coord_min for f in frames[:nd]]),
                              enc.frame_bytes[:nd], enc.point_nums[:nd])
    pipeline.decode_gop(sub, cx.dev)                      # warm-up (allocations, streams)
    best = float("inf")
    for _ in range(2):                                    # host-thread scheduling makes single runs noisy: best of two
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dec = pipeline.decode_gop(sub, cx.dev)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / nd)
    lossless = all(bool(d.shape == p.shape and (d == p).all()) for d, p in zip(dec, pts_dev[:nd]))
    return lossless, best, sub


def run_job(cx: Ctx, args, shape, G, K, W, e2e=True, with_checks=True):
    """Headline measure: the whole sequence.  Returns the JSON line (dict) on every rank (rank 0's is printed)."""
    F, E = args.frames, args.epochs
    job = Job(cx, shape, F, G, E)
    n_frames = F * G
    breakdown = []
    for i in range(W):
        if i == W - 1:
            cx.timed(job.step, 1, prof_mask=cx.all_mask(), serial=True)
            breakdown = cx.prof_table()
        else:
            job.step()
    dom = max(range(len(breakdown)), key=lambda c: breakdown[c]["ms"]) if breakdown else 0
    clocks = ClockSampler(cx.local)
    if cx.rank == 0:
        clocks.start()
    ms, wall, enc = cx.timed(job.step, K, prof_mask=1 << dom)
    live = cx.prof_table()
    clk = clocks.stop() if cx.rank == 0 else None
    launches = sum(r["launches"] for r in live)
    first_frames = job.frames[job.mine[0][0]]
    pbar = mean_occupied(first_frames)
    line = {"metric": METRIC, "value": ms / 1e3 / K / n_frames, "unit": UNIT, "n_gpus": cx.world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": None, "clocks": clk, "e2e": None, "gpu_launches": launches,
            "roofline": add_alone(roofline_of(live[dom], pbar, ms), breakdown[dom], pbar) if live else None,
            "roofline_step": roofline_step(live, pbar, ms) if live else None,
            "kernel_breakdown_ms_per_step": {r["kernel"]: round(r["ms"], 3) for r in breakdown if r["launches"]},
            "kernel_breakdown_note": "one fully profiled warm-up step with every launch on one stream; the timed steps overlap the "
                                     "weight-gradient launches with the grad-input chain on a second stream and are shorter than this sum",
            "wall_s_timed": wall, "host_cores_per_rank": cx.cores or (os.cpu_count() or 0)}
    if e2e:
        for _ in range(min(W, 1)):
            job.step(from_host=True)
        Ke = min(K, 3)
        ms_e, _, _ = cx.timed(lambda: job.step(from_host=True), Ke)
        ms_e *= K / Ke   # scaled to K steps below; the e2e region is timed over min(K, 3) whole jobs
        io = torch.tensor([job.h2d_bytes(), job.d2h_bytes()], dtype=torch.float64, device=cx.dev)
        if cx.world > 1:
            cx.dist.all_reduce(io)
        line["e2e"] = {"value": ms_e / 1e3 / K / n_frames, "unit": UNIT, "h2d_bytes_per_step": int(io[0].item()),
                       "d2h_bytes_per_step": int(io[1].item()), "steps": Ke}
        # what the e2e - resident difference is made of: per-frame upload + octree / kernel-map / pair-list preparation
        tabs = first_frames[0].tables
        prep_bytes = sum(int(t.numel()) * t.element_size() for t in (tabs.coords, tabs.anchor, tabs.mask, tabs.nbr7, tabs.occ, tabs.scale,
                                                                        tabs.tile_rng, tabs.pair_cnt, tabs.pair_list) if t is not None)
        d_ms = (ms_e - ms) / K / n_frames
        line["roofline_prep"] = {"bound": "hbm", "stage": "frame preparation (H2D + sort/unique + hash + kernel map + pair lists)",
                                 "ms_per_frame": d_ms, "table_bytes_per_frame": prep_bytes,
                                 "achieved": prep_bytes / max(d_ms, 1e-9) / 1e6, "unit": "GB/s", "peak": line["roofline"]["peak"] if line["roofline"] else None,
                                 "frac": (prep_bytes / max(d_ms, 1e-9) / 1e6) / line["roofline"]["peak"] if line["roofline"] else None,
                                 "note": "table bytes written per frame / (e2e - resident) time; sort passes and hash probes re-read them several times; "
                                         "only a rank's first GOP is prepared in line, the later ones in the background while the GOP "
                                         "before them is overfitted (pipeline.GopPreparer)"}
    # phase timings (one extra untimed-for-the-headline pass each): GOP 0's fit, whole-job fit, encode
    if with_checks:
        ms_fit, _, _ = cx.timed(lambda: job.step(encode=False), 1)
        line["overfit_iters_per_s"] = n_frames * E / (ms_fit / 1e3)          # frame-iterations (fwd+bwd+Adam) per second, whole job
        line["encode_s_per_frame"] = max(0.0, (ms / K - ms_fit)) / 1e3 / n_frames
        g0, ranks0 = job.mine[0]
        tr0 = job.trainers[tuple(ranks0)]
        tr0.reset(seed=8807)
        ms_g0, _, _ = cx.timed(lambda: tr0.fit(job.frames[g0], 1), 1)
        line["gop0_iters_per_s"] = F / (ms_g0 / 1e3)                          # GOP 0: every rank on the same frame
        line["gop0_split"] = len(ranks0)
    if cx.rank == 0 and enc:
        line["bpp"] = {f"gop{g}": e.bpp for g, e in sorted(enc.items())}
        line["points_per_frame"] = int(np.mean(enc[0].point_nums))
        line["voxel_passes_per_frame"] = int(np.mean([f.tables.n_rows for f in first_frames]))
        if with_checks:
            lossless, dec_s, _ = decode_check(cx, job.pipeline, enc[0], job.frames[0], job.pts_dev[0], job.S)
            line["decode_lossless"], line["decode_s_per_frame"] = lossless, dec_s
    return line, job


def run_replica(cx: Ctx, args, K, W):
    """Round-1 measure: one independent GOP per GPU, no collective (weak scaling)."""
    from linr_pcgc_b200 import pipeline, synth
    from linr_pcgc_b200.net import NetRunner
    from linr_pcgc_b200.trainer import GopTrainer
    F, E = args.frames, args.epochs
    pts = synth.make_sequence(args.shape, F, device=cx.dev, start=cx.rank * F)
    frames = pipeline.prepare_gop(pts, None, 64, cx.dev)
    S = frames[0].n_scales
    mr = max(f.tables.n_rows for f in frames)
    tr = GopTrainer(S, cx.dev, seed=8807, max_rows=mr)
    run = NetRunner(S, mr, cx.dev, train=False)

    def step():
        tr.fit(frames, E)
        return pipeline.encode_gop(frames, tr.state.params, S, 8, runner=run)

    breakdown = []
    for i in range(W):
        if i == W - 1:
            cx.timed(step, 1, prof_mask=cx.all_mask(), serial=True)
            breakdown = cx.prof_table()
        else:
            step()
    ms, wall, enc = cx.timed(step, K)
    return {"value": ms / 1e3 / K / (F * cx.world), "unit": UNIT, "scaling": "weak", "steps": K, "ms_per_step": ms / K,
            "frames_per_step": F * cx.world, "parallelism": "gop%d (one independent GOP per GPU, no collective)" % cx.world,
            "bpp": enc.bpp, "kernel_breakdown_ms_per_step": {r["kernel"]: round(r["ms"], 3) for r in breakdown if r["launches"]}}, (frames, pts, enc, S, tr, run)


def run_decode_only(cx: Ctx, args, shape, F=16, E=3, K=3, W=1):
    """BASELINE.json configs[4]: decode-only throughput (deterministic inference + host arithmetic decoding) of an
    MVUB-shaped GOP, bit-exact reconstruction checked.  The GOP is overfitted for a few epochs and encoded outside the
    timed region; a step decodes all its frames (codec.decode_frames_batched: lockstep batches of up to 8 frames -- one
    launch set and one device<->host round trip per stage for the whole batch -- up to four batches in flight)."""
    from linr_pcgc_b200 import pipeline, synth
    from linr_pcgc_b200.trainer import GopTrainer
    pts = synth.make_sequence(shape, F, device=cx.dev, start=cx.rank * F)
    frames = pipeline.prepare_gop(pts, None, 64, cx.dev)
    S = frames[0].n_scales
    tr = GopTrainer(S, cx.dev, seed=8807, max_rows=max(f.tables.n_rows for f in frames))
    tr.fit(frames, E)
    enc = pipeline.encode_gop(frames, tr.state.params, S, 8)
    for _ in range(W):
        pipeline.decode_gop(enc, cx.dev)
    # kernel time of one decode of the GOP, from a separate fully profiled pass: the per-launch events (and their mutex,
    # taken by every decoder thread) would slow the timed passes down
    cx.lib.linr_prof_enable(cx.all_mask())
    pipeline.decode_gop(enc, cx.dev)
    torch.cuda.synchronize()
    tab = cx.prof_table()
    cx.lib.linr_prof_enable(0)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        dec = pipeline.decode_gop(enc, cx.dev)
    cx.barrier()
    wall = time.perf_counter() - t0          # the decoder's streams belong to its worker threads: host clock around a full sync
    t = torch.tensor([wall], dtype=torch.float64, device=cx.dev)
    if cx.world > 1:
        cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
    wall = float(t.item())
    lossless = all(bool(d.shape == p.shape and (d == p).all()) for d, p in zip(dec, pts))
    rows = float(np.mean([f.tables.n_rows for f in frames]))
    s_frame = wall / K / (F * cx.world)
    conv_ms = sum(r["ms"] for r in tab if r["kernel"].startswith("conv27"))
    fwd_bytes = 7950.0 * rows                                   # forward-only algorithmic bytes per voxel-pass
    peak = roofline_of({"kernel": "conv27<8,8>", "units": 0, "ms": 1, "launches": 0}, 14.3, 1.0)["peak"]
    return {"metric": "decode_s_per_frame", "value": s_frame, "unit": UNIT, "shape": shape, "frames": F * cx.world, "steps": K,
            "lossless": lossless, "bpp": enc.bpp, "points_per_frame": int(np.mean(enc.point_nums)), "voxel_passes_per_frame": int(rows),
            "schedule": "lockstep batches of up to 8 frames, up to 4 batches in flight (codec.decode_frames_batched)", "gpu_launches": sum(r["launches"] for r in tab) * K,
            "roofline": {"bound": "hbm", "kernel": "sequential 8-stage inference (all conv27 classes)", "unit": "GB/s",
                         "achieved_wall": fwd_bytes / s_frame / 1e9, "achieved_kernels": fwd_bytes * F / max(conv_ms, 1e-9) / 1e6,
                         "peak": peak, "frac_wall": fwd_bytes / s_frame / 1e9 / peak,
                         "frac_kernels": fwd_bytes * F / max(conv_ms, 1e-9) / 1e6 / peak,
                         "note": "wall: host range decoder + 56 device<->host round trips per batch included; kernels: CUDA-event time of the conv launches only"}}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    cx = Ctx(args)
    W, K = max(args.warmup, 0), max(args.steps, 1)
    G = args.gops or (2 if args.shape == "owlii" else 3)

    if args.decode_only:
        line = run_decode_only(cx, args, args.shape, F=min(args.frames, 16), E=min(args.epochs, 3), K=K, W=max(1, min(W, 2)))
        line.update({"n_gpus": cx.world, "higher_is_better": False, "dtype": "f32", "data": "synthetic",
                     "config": {"workload": f"{args.shape}-shaped decode-only (BASELINE.json configs[4])"}})
        if cx.rank == 0:
            print(json.dumps(line), flush=True)
        return finish(cx)

    if args.mode == "gop":
        rep, (frames, pts, enc, S, tr, run) = run_replica(cx, args, K, W)
        line = {"metric": METRIC, "value": rep["value"], "unit": UNIT, "n_gpus": cx.world, "steps": K, "warmup": W,
                "ms_per_step": rep["ms_per_step"], "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args, cx.world, G), "bpp": rep["bpp"],
                "kernel_breakdown_ms_per_step": rep["kernel_breakdown_ms_per_step"]}
        if cx.rank == 0:
            print(json.dumps(line), flush=True)
        return finish(cx)

    line, job = run_job(cx, args, args.shape, G, K, W, e2e=not args.no_e2e)
    line["config"] = workload_config(args, cx.world, G)
    if not args.no_extras:
        rep, _ = run_replica(cx, args, 1, 1)
        line["replica"] = rep
        extras = {}
        try:
            if cx.world == 1 and args.shape == "loot":
                del job
                torch.cuda.empty_cache()
                for shp in ("mvub10", "mvub9"):      # C5: decode-only throughput
                    extras[f"C5_decode_{shp}"] = run_decode_only(cx, args, shp, F=16, E=3, K=2, W=1)
            if (cx.world == 8 or os.environ.get("LINR_BENCH_C4") == "1") and args.shape == "loot":
                del job
                torch.cuda.empty_cache()
                a2 = argparse.Namespace(**vars(args))
                a2.shape = "owlii"
                l2, j2 = run_job(cx, a2, "owlii", 2, 1, 1, e2e=False, with_checks=False)   # C4: Owlii 64 frames on 8 GPUs
                l2["config"] = workload_config(a2, cx.world, 2)
                extras["C4_owlii_64_frames"] = {k: l2[k] for k in ("value", "unit", "ms_per_step", "steps", "scaling", "config", "bpp",
                                                                 "points_per_frame", "voxel_passes_per_frame", "roofline") if k in l2}
        except Exception as ex:   # side configs never take the headline down
            extras["error"] = repr(ex)
        line["side_configs"] = extras
    if cx.rank == 0 and cx.world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_time(args, 1, 0, full=args.cpu_full_frame)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"]}
        except Exception as ex:  # the baseline is reported, never load-bearing
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    if cx.rank == 0:
        print(json.dumps(line), flush=True)
    finish(cx)


def finish(cx):
    if cx.world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per (row x group) unit of every kernel class, from ONE `ncu --set full`
# capture of a training iteration, written by tools/ncu_traffic_table.py (profiles/r02_traffic_per_unit.json); per unit,
# so it scales to the launches of this run.  A class the capture does not hold has no entry (traffic: null).
def _load_traffic():
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_per_unit.json")))
        return ({k: v["dram_bytes_per_unit"] for k, v in t["classes"].items()},
                f"ncu dram__bytes_read/write.sum over one frame-iteration ({t['source']}, {t['rows_per_frame']} rows/frame), tools/ncu_traffic_table.py -> "
                "profiles/r02_traffic_per_unit.json; per (row x group) unit, scaled by this run's units per launch")
    except Exception:
        return {}, "profiles/r02_traffic_per_unit.json missing"


NCU_DRAM_BYTES_PER_UNIT, NCU_TRAFFIC_SOURCE = _load_traffic()


def algorithmic_bytes(kernel: str, pbar: float) -> float:
    """Algorithmic bytes per (row x group) unit of a kernel class (DESIGN.md 'Roofline model'):
    fp32 features read once + written once, kernel map = 4 B per occupied neighbour + 4 B mask, weights free."""
    km = 4.0 * pbar + 4.0
    table = {"conv27<8,8>": 4 * (8 + 8) + km, "conv27<8,4>": 4 * (8 + 4) + km, "conv27<4,8>": 4 * (4 + 8) + km,
             "conv27<4,4>": 4 * (4 + 4) + km, "conv27_bits<8>": 1 + 4 * 8 + km, "conv27_head": 4 * 8 + km + 2 + 1,
             "bwd_w<8,8>": 4 * (8 + 8) + km, "bwd_w<8,4>": 4 * (8 + 4) + km, "bwd_w<4,4>": 4 * (4 + 4) + km,
             "bwd_w_bits<8>": 1 + 4 * 8 + km, "pointwise": 4 * 12, "pointwise_bwd_w": 4 * 12, "head_bwd": 4 * (8 + 8 + 1),
             "sce": 1 + 4 * 8, "reduce": 4 * 9, "adam_quant": 4 * 7, "coord": 12}
    return table.get(kernel, 0.0)


if __name__ == "__main__":
    main()
