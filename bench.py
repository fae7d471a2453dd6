#!/usr/bin/env python
"""bench.py — overfit-plus-encode seconds per frame of the LINR-PCGC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A STEP is one GOP of the loot-shaped workload (BASELINE.json configs[1]): `--epochs` (10) passes of per-frame
forward + backward + Adam over `--frames` (32) synthetic 10-bit frames of ~780k points, then 8-bit model
quantisation and the real encode of every frame (network forward, 16-bit CDFs to the host, range coder).
  value : inputs (prepared frames) resident in HBM when the timed region starts.
  e2e   : the public API `pipeline.overfit_encode_gop` fed pinned HOST point arrays: H2D copy, octree/kernel-map
          preparation, overfit, encode, bitstreams back on the host — all inside the timed region.
N > 1 (torchrun): independent GOPs, one per GPU, no collective on the data path (weak scaling); `--dp` instead splits
the frames of ONE GOP across ranks with an NCCL all-reduce of the 219 kB gradient per optimiser step.
`--impl reference` times the CPU restatement of the reference (oracle/, torch CPU, all host threads) on a bounded
sample of the same workload.
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
    ap.add_argument("--epochs", type=int, default=10, help="first_epoch / others_epoch of the north-star config")
    ap.add_argument("--dp", action="store_true", help="intra-GOP data parallel instead of one GOP per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gop-pipeline", action="store_true", help="code GOP i in the background while GOP i+1 is overfitted")
    ap.add_argument("--cpu-sample-rows", type=int, default=75000)
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
    """A bounded sample of one frame of the workload for the CPU restatement: the coarsest scales of one frame whose
    parent-voxel count stays below `sample_rows` (cost is linear in voxel-passes; the ratio is reported)."""
    from linr_pcgc_b200 import params as P, synth
    from oracle import linr_oracle as O
    torch.set_num_threads(threads)
    pts = synth.make_sequence(shape, 1)[0].numpy()
    fr = O.prepare_frame(pts, None, 64)
    rows = [len(s["coord"]) for s in fr["scales"]]
    total = sum(rows)
    keep, acc = [], 0
    for i in range(len(rows) - 1, -1, -1):
        if acc + rows[i] > sample_rows and keep:
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


def cpu_reference_time(args, steps: int, warmup: int):
    threads = os.cpu_count() or 1
    O, sub, nbrs, S, flat, rows, total, point_num = oracle_sample(args.shape, args.cpu_sample_rows, threads)
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
    t_enc = time.perf_counter() - t0
    t_iter = float(np.mean(its)) * scale
    t_enc *= scale
    s_per_frame = args.epochs * t_iter + t_enc
    sample = (f"oracle port (torch CPU), {rows} of {total} voxel-passes of one {args.shape} frame (coarsest scales), "
              f"{steps} frame-iterations + 1 encode timed, scaled by {scale:.2f} to the full frame; "
              f"s/frame = {args.epochs} x iter + encode")
    return s_per_frame, t_iter, t_enc, threads, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    steps = min(steps, 3)  # each step is ~10 s of CPU work on the bounded sample
    t_all0 = time.perf_counter()
    v, t_iter, t_enc, threads, sample = cpu_reference_time(args, steps, warmup)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": v * args.frames * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "iter_s": t_iter, "encode_s": t_enc, "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    from linr_pcgc_b200 import synth
    bits, pts = synth.SHAPES[args.shape]
    cfg = {"loot": "BASELINE.json configs[1]", "owlii": "BASELINE.json configs[3] (per-GPU share)",
           "plumbing": "BASELINE.json configs[0]"}.get(args.shape, "parity-test shape")
    return {"workload": f"{args.shape}-shaped synthetic {bits}-bit surface, ~{pts // 1000}k pts/frame, gop_size {args.frames}, "
                        f"{args.epochs} epochs/GOP, overfit + model quantisation + encode ({cfg})",
            "gop_size": args.frames, "epochs": args.epochs, "frames_per_step": args.frames * (1 if args.dp else n),
            "parallelism": ("dp%d (frames of one GOP split, NCCL all-reduce of gradients)" % n) if args.dp and n > 1
            else ("gop%d (one GOP per GPU, no collective)" % n),
            "gop_pipeline": "coding of GOP i overlaps overfitting of GOP i+1 (all K GOPs coded inside the timed region)" if args.gop_pipeline else "serial",
            "l2_policy": "inputs larger than L2: one GOP's resident tables + activations >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------ B200 arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from linr_pcgc_b200 import _lib, pipeline, synth
    from linr_pcgc_b200.net import NetRunner
    from linr_pcgc_b200.trainer import GopTrainer
    lib = _lib.load()

    F, E = args.frames, args.epochs
    if args.dp and world > 1:
        my = list(range(rank, F, world))       # frames of ONE GOP split across ranks
        start = 0
    else:
        my = list(range(F))                    # one GOP per rank
        start = rank * F
    seq = synth.make_sequence(args.shape, F, device=dev, start=start)
    pts_dev = [seq[i] for i in my]
    pts_host = [p.cpu().pin_memory() for p in pts_dev]
    frames = pipeline.prepare_gop(pts_dev, None, 64, dev)
    S = frames[0].n_scales
    rows = [f.tables.n_rows for f in frames]
    max_rows = max(rows)

    grad_hook = None
    if args.dp and world > 1:
        def grad_hook(g):
            dist.all_reduce(g)   # sum of per-frame gradients; every rank then takes the same Adam step
    tr = GopTrainer(S, dev, seed=8807, max_rows=max_rows, grad_hook=grad_hook)
    run = NetRunner(S, max_rows, dev, train=False)

    # --gop-pipeline: the coding of the GOP of step i (side stream + host coder threads) overlaps the overfitting of
    # the GOP of step i+1; `timed` collects the last one before it closes the timed region, so K steps = K GOPs fully
    # overfitted AND coded.  Measured round 1: -1 % on resident inputs, +4 % end to end (the coder's host threads
    # compete with the launch thread) -> off by default, every step is strictly serial.
    coder = pipeline.GopCoder(dev) if args.gop_pipeline else None

    def step_resident():
        tr.fit(frames, E)
        if coder is not None:
            return coder.submit(frames, tr.state.params, S)
        return pipeline.encode_gop(frames, tr.state.params, S, 8, runner=run)

    state = {"s": None}

    def step_e2e():
        enc, st, _ = pipeline.overfit_encode_gop(pts_host, E, state=state["s"], device=dev, seed=8807,
                                                 trainer_kwargs={"grad_hook": grad_hook}, coder=coder)
        state["s"] = st
        return enc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, prof_mask=0):
        lib.linr_prof_enable(prof_mask)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        if coder is not None and hasattr(out, "result"):
            out = coder.collect()     # the last GOP's bitstreams; orders this stream after the coder's
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, out

    def prof_table():
        import ctypes as C
        tab = []
        for c in range(lib.linr_prof_classes()):
            ms, n, u = C.c_double(), C.c_int64(), C.c_int64()
            lib.linr_prof_read(c, C.byref(ms), C.byref(n), C.byref(u))
            tab.append({"kernel": lib.linr_prof_name(c).decode(), "ms": ms.value, "launches": n.value, "units": u.value})
        return tab

    # warm-up: W untimed steps; the last one runs with every kernel class bracketed by events -> breakdown
    W, K = max(args.warmup, 0), max(args.steps, 1)
    for i in range(W):
        if i == W - 1:
            _, _, _ = timed(step_resident, 1, prof_mask=(1 << lib.linr_prof_classes()) - 1)
            breakdown = prof_table()
        else:
            step_resident()
            if coder is not None:
                coder.collect()
    if W == 0:
        breakdown = []
    dom = max(range(len(breakdown)), key=lambda c: breakdown[c]["ms"]) if breakdown else 0

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, wall, enc = timed(step_resident, K, prof_mask=1 << dom)
    live = prof_table()
    clk = clocks.stop() if rank == 0 else None
    launches = sum(r["launches"] for r in live)

    # roofline of the dominant kernel class, measured live in the timed region
    mask_pop = 0.0
    for f in frames[:4]:
        m = f.tables.mask.to(torch.int64) & 0x7FFFFFF
        cnt = torch.zeros_like(m)
        for b in range(27):
            cnt += (m >> b) & 1
        mask_pop += float(cnt.double().mean().item())
    pbar = mask_pop / max(1, len(frames[:4]))
    d = live[dom]
    bytes_per_unit = algorithmic_bytes(d["kernel"], pbar)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = (bytes_per_unit * d["units"]) / max(d["ms"], 1e-9) / 1e6 if d["launches"] else 0.0  # GB/s
    roofline = {"bound": "hbm", "kernel": d["kernel"], "achieved": achieved, "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": (NCU_DRAM_BYTES_PER_UNIT[d["kernel"]] * d["units"] / max(1, d["launches"])
                            if d["kernel"] in NCU_DRAM_BYTES_PER_UNIT else None),
                "traffic_source": "ncu --set full dram bytes per (row x group) unit, profiles/r01e_*, r01g_*; scaled by this run's units per launch",
                "launches": d["launches"], "avg_launch_us": 1e3 * d["ms"] / max(1, d["launches"]),
                "algorithmic_bytes_per_launch": bytes_per_unit * d["units"] / max(1, d["launches"]),
                "mean_occupied_neighbours": pbar,
                "share_of_step": d["ms"] / max(ms, 1e-9)}

    # end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        for _ in range(min(W, 1)):
            step_e2e()
            if coder is not None:
                coder.collect()
        ms_e, _, enc_e = timed(step_e2e, K)
        n_frames_job = F if (args.dp and world > 1) else F * world
        h2d = sum(int(p.numel()) * 4 for p in pts_host)
        d2h = sum(r * (8 * 2 + 1) for r in rows) + 8 * len(rows) * E + 54712
        e2e = {"value": ms_e / 1e3 / K / n_frames_job, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    # decode the whole GOP outside the headline region: lossless check (decoder.py:140) + decode throughput
    lossless, decode_s = None, None
    if rank == 0:
        nd = min(len(frames), 16)
        sub = pipeline.EncodedGop(enc.scale_num, enc.side_info, enc.model_bytes, enc.model_bits,
                                  pipeline.codec.pack_low_xyz([f.scale_coords(S - 1).cpu().numpy() for f in frames[:nd]],
                                                              [f.coord_min for f in frames[:nd]]),
                                  enc.frame_bytes[:nd], enc.point_nums[:nd])
        pipeline.decode_gop(sub, dev)                      # warm-up (allocations, streams)
        decode_s = float("inf")
        for _ in range(2):                                 # host-thread scheduling makes single runs noisy: best of two
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dec = pipeline.decode_gop(sub, dev)
            torch.cuda.synchronize()
            decode_s = min(decode_s, (time.perf_counter() - t0) / nd)
        lossless = all(bool(d.shape == p.shape and (d == p).all()) for d, p in zip(dec, pts_dev[:nd]))

    n_frames_job = F if (args.dp and world > 1) else F * world
    value = ms / 1e3 / K / n_frames_job
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": False, "scaling": "strong" if (args.dp and world > 1) else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "overfit_iters_per_s": None, "bpp": enc.bpp, "decode_lossless": lossless, "decode_s_per_frame": decode_s,
            "points_per_frame": int(np.mean(enc.point_nums)), "voxel_passes_per_frame": int(np.mean(rows)),
            "kernel_breakdown_ms_per_step": {r["kernel"]: round(r["ms"], 3) for r in breakdown if r["launches"]},
            "wall_s_timed": wall}

    # overfit-only and encode-only split (one extra untimed-for-the-headline pass each)
    ms_fit, _, _ = timed(lambda: tr.fit(frames, 1), 1)
    ms_enc, _, _ = timed(lambda: pipeline.encode_gop(frames, tr.state.params, S, 8, runner=run), 1)
    line["overfit_iters_per_s"] = n_frames_job / (ms_fit / 1e3)   # frame-iterations (fwd+bwd+Adam) per second, whole job
    line["encode_s_per_frame"] = ms_enc / 1e3 / n_frames_job

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, t_iter, t_enc, threads, sample = cpu_reference_time(args, 1, 0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        except Exception as ex:  # the baseline is reported, never load-bearing
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE 8-group launch over a loot-shaped frame (277,011 rows x 8 groups
# = 2,216,088 units), from the `ncu --set full` captures summarised under profiles/ (r01e conv kernels, r01g weight
# gradients); per unit, so it scales to the launches of this run.  None: not captured.
NCU_DRAM_BYTES_PER_UNIT = {
    "conv27<8,8>": (83.960832e6 + 39.300096e6) / 2216088, "conv27<8,4>": (82.984704e6 + 18.523136e6) / 2216088,
    "conv27<4,4>": (173.155072e6 + 58.067968e6) / 2216088, "conv27_bits<8>": (11.316992e6 + 6.055936e6) / 1939077,
    "conv27_head": (84.780544e6 + 47.52e6) / 2216088, "bwd_w<8,8>": (153.3184e6 + 8.76288e6) / 2216088,
    "bwd_w<4,4>": (118.37824e6 + 4.628224e6) / 2216088, "bwd_w_bits<8>": (73.083136e6 + 3.748608e6) / 1939077,
}


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
