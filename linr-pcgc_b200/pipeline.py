"""GOP-level public API: prepare -> overfit -> quantise model -> encode (-> decode), one call per GOP.

Host-side mirror of the reference's orchestration for one GOP: `overfit_one_gop` (main.py:122-455),
`encode_one_gop` (encoder.py:57-156) and `decode_one_gop` (decoder.py:51-147), minus the file IO (callers that want
the reference's on-disk layout use `write_gop` / `read_gop`).  Everything a frame needs stays resident in HBM
between epochs (the reference re-reads a pickle per iteration, datautils/custom_dataset.py:235-241).
"""
from __future__ import annotations

import dataclasses
import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import codec, model_compression
from . import params as P
from .frame import Frame, prepare_frame
from .net import NetRunner
from .trainer import GopTrainer, OptimState


@dataclass
class EncodedGop:
    """What the reference writes under <encode_dir>/<gop>/ (encoder.py:84-146)."""
    scale_num: int
    side_info: Dict                      # mu, b, min_param, max_param, enc_mode, bitdepth  -> side_info.json
    model_bytes: bytes                   # bins/model.bin
    model_bits: float                    # bit_real of the model coder (model_size_est.py:489)
    low_enc_bytes: bytes                 # bins/low_enc_bytes.bin
    frame_bytes: List[List[bytes]]       # bins/frame%04d_scale%d.bin
    point_nums: List[int]

    @property
    def total_bits(self) -> float:
        return sum(8 * len(b) for fb in self.frame_bytes for b in fb) + self.model_bits + 8 * len(self.low_enc_bytes)

    @property
    def bpp(self) -> float:
        """`real_bpp_all` accounting of test_utils.py:145-157."""
        return self.total_bits / max(1, sum(self.point_nums))


def prepare_gop(points: Sequence[torch.Tensor], scale_num: Optional[int] = None, min_point_num: int = 64,
                device="cuda") -> List[Frame]:
    """Upload (if on the host) and prepare every frame of a GOP.  `scale_num` None: discovered from the first frame
    and then caps the others (main.py:77-78)."""
    frames: List[Frame] = []
    for p in points:
        if not p.is_cuda:
            p = p.to(device, non_blocking=True)
        f = prepare_frame(p, scale_num, min_point_num)
        if scale_num is None:
            scale_num = f.n_scales
        frames.append(f)
    return frames


def _side_info(comp: Dict, bitdepth: int) -> Dict:
    """side_info.json (encoder.py:114); `cdf_version` only appears when it is not the reference's (so version-1 GOPs
    stay byte-compatible with the reference's decoder)."""
    side = dict(mu=comp["mu"], b=comp["b"], min_param=comp["min_param"], max_param=comp["max_param"],
                enc_mode=comp["enc_mode"], bitdepth=bitdepth)
    if int(comp.get("cdf_version", 1)) != model_compression.CDF_REFERENCE:
        side["cdf_version"] = int(comp["cdf_version"])
    return side


def encode_gop(frames: Sequence[Frame], flat_params: torch.Tensor, scale_num: int, bitdepth: int = 8,
               runner: Optional[NetRunner] = None, threads: Optional[int] = None,
               cdf_version: int = model_compression.CDF_REFERENCE) -> EncodedGop:
    """Quantise the model, then code every frame with the *dequantised* parameters (encoder.py:101-103)."""
    comp = model_compression.compress_model(flat_params, bitdepth, cdf_version)
    recon = comp["recon_ret"]
    if runner is None:
        runner = NetRunner(scale_num, max(f.tables.n_rows for f in frames), flat_params.device, train=False)
    frame_bytes = codec.encode_frames(runner, recon, frames, threads)
    lows = [f.scale_coords(f.n_scales - 1).cpu().numpy() for f in frames]
    low = codec.pack_low_xyz(lows, [f.coord_min for f in frames])
    return EncodedGop(scale_num, _side_info(comp, bitdepth), comp["final_bytes"], comp["bit_real"], low, frame_bytes,
                      [f.point_num for f in frames])


def encode_gop_shared(frames: Sequence[Frame], flat_params: torch.Tensor, scale_num: int, bitdepth: int = 8,
                      runner: Optional[NetRunner] = None, ranks: Optional[Sequence[int]] = None, group=None) -> Optional[EncodedGop]:
    """encode_gop with the frames of the GOP dealt to the `ranks` that trained it together (every member holds the same
    parameters bit for bit after a stage-split fit): member p codes frames p, p + parts, ...; the first rank collects the
    bitstreams and returns the EncodedGop, the others return None."""
    import torch.distributed as dist
    if not ranks or len(ranks) == 1:
        return encode_gop(frames, flat_params, scale_num, bitdepth, runner=runner)
    part = list(ranks).index(dist.get_rank())
    comp = model_compression.compress_model(flat_params, bitdepth)
    if runner is None:
        runner = NetRunner(scale_num, max(f.tables.n_rows for f in frames), flat_params.device, train=False)
    mine = list(range(part, len(frames), len(ranks)))
    coded = codec.encode_frames(runner, comp["recon_ret"], [frames[i] for i in mine])
    box = [None] * len(ranks) if part == 0 else None
    dist.gather_object(list(zip(mine, coded)), box, dst=ranks[0], group=group)
    if part != 0:
        return None
    frame_bytes: List = [None] * len(frames)
    for share in box:
        for i, fb in share:
            frame_bytes[i] = fb
    lows = [f.scale_coords(f.n_scales - 1).cpu().numpy() for f in frames]
    low = codec.pack_low_xyz(lows, [f.coord_min for f in frames])
    return EncodedGop(scale_num, _side_info(comp, bitdepth), comp["final_bytes"], comp["bit_real"], low, frame_bytes,
                      [f.point_num for f in frames])


def decode_gop(enc: EncodedGop, device="cuda", workers: Optional[int] = None, batched: bool = True) -> List[torch.Tensor]:
    """decode_one_gop (decoder.py:51-147): model from its bitstream, frames coarse-to-fine; returns the original
    (min-restored) sorted coordinates of every frame as CUDA int32 [Np,3]."""
    n = P.offsets(P.param_spec(enc.scale_num))[-1]
    d = dict(enc.side_info)
    d["final_bytes"] = enc.model_bytes
    flat = model_compression.decompress_model(d, n, device)
    lows, mins = codec.unpack_low_xyz(enc.low_enc_bytes)
    jobs = [(fb, torch.from_numpy(lows[i]).to(device)) for i, fb in enumerate(enc.frame_bytes)]
    if batched and len(jobs) >= 2:
        dec = codec.decode_frames_batched(flat, enc.scale_num, jobs)
    else:
        dec = codec.decode_frames(flat, enc.scale_num, jobs, workers=workers or min(16, codec.rc.host_cores()))
    return [xyz + torch.from_numpy(mins[i].copy()).to(device) for i, xyz in enumerate(dec)]


class GopPreparer:
    """Uploads and prepares GOP g+1 (prepare_gop: H2D, octree levels, hash, kernel map, pair lists) in the background
    while the caller overfits GOP g.

    The preparation of a frame is a chain of small kernels with a host read of a row count between the levels: it leaves
    the GPU mostly idle and costs the caller 1.0-1.3 ms per frame when done in line.  Here it runs on its own stream and
    host thread; its host reads wait for that stream only, and the caller's launch thread stays far enough ahead of the GPU
    for the two to share the interpreter.  One GOP is in flight at a time."""

    def __init__(self, device="cuda"):
        from concurrent.futures import ThreadPoolExecutor
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.side = torch.cuda.Stream(self.device)
        self.pool = ThreadPoolExecutor(max_workers=1)
        self.pending = None

    def submit(self, points: Sequence[torch.Tensor], scale_num: Optional[int] = None, min_point_num: int = 64):
        assert self.pending is None, "one GOP in flight at a time: collect() first"
        ready = torch.cuda.Event()
        ready.record()                       # device-resident points may still be in the making on the caller's stream

        def work():
            torch.cuda.set_device(self.device)
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                frames = prepare_gop(points, scale_num, min_point_num, self.device)
                self.side.synchronize()
            return frames

        self.pending = self.pool.submit(work)

    def collect(self) -> List[Frame]:
        """The prepared frames, usable on the caller's current stream."""
        frames = self.pending.result()
        self.pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.side)
        for f in frames:                     # allocated on the side stream, used (and outlived) on the caller's
            for t in (f.xyz, *(getattr(f.tables, k.name) for k in dataclasses.fields(f.tables))):
                if isinstance(t, torch.Tensor):
                    t.record_stream(cur)
        return frames


class GopCoder:
    """Codes GOP g in the background while the caller overfits GOP g+1.

    The coding of a GOP needs only its frames (read-only tables) and a snapshot of the trained parameters, and every
    GOP after the first starts from GOP 0's state (main.py:102-104), so consecutive GOPs pipeline: the network
    forward of the coder runs on a side stream, the CDF download and the range coder on host threads, the next GOP's
    overfitting on the caller's stream.  One GOP is in flight at a time (staging buffers and the inference workspace
    are reused); `submit` of the next GOP first collects the previous one."""

    def __init__(self, device="cuda", bitdepth: int = 8, threads: Optional[int] = None):
        from concurrent.futures import ThreadPoolExecutor
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.bitdepth, self.threads = bitdepth, threads
        self.side = torch.cuda.Stream(self.device)
        self.pool = ThreadPoolExecutor(max_workers=1)
        self.pending = None
        self.runner: Optional[NetRunner] = None
        self.results: List[EncodedGop] = []

    def submit(self, frames: Sequence[Frame], flat_params: torch.Tensor, scale_num: int):
        """Call on the stream that trained `flat_params`; returns a future of the EncodedGop."""
        self.collect()
        snap = flat_params.clone()            # the caller keeps training this vector
        snap.record_stream(self.side)
        ready = torch.cuda.Event()
        ready.record()
        if self.runner is None or self.runner.S != scale_num:
            self.runner = NetRunner(scale_num, max(f.tables.n_rows for f in frames), self.device, train=False)
        runner = self.runner

        def work():
            torch.cuda.set_device(self.device)
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                enc = encode_gop(frames, snap, scale_num, self.bitdepth, runner=runner, threads=self.threads)
                self.side.synchronize()
            return enc

        self.pending = self.pool.submit(work)
        return self.pending

    def collect(self) -> Optional[EncodedGop]:
        """Wait for the GOP in flight (if any) and order the caller's stream after the coder's."""
        if self.pending is None:
            return None
        enc = self.pending.result()
        self.pending = None
        torch.cuda.current_stream(self.device).wait_stream(self.side)
        self.results.append(enc)
        return enc


def overfit_encode_gop(points: Sequence[torch.Tensor], epochs: int, state: Optional[OptimState] = None,
                       scale_num: Optional[int] = None, min_point_num: int = 64, bitdepth: int = 8, device="cuda",
                       seed: Optional[int] = None, trainer_kwargs: Optional[Dict] = None, threads: Optional[int] = None,
                       coder: Optional[GopCoder] = None):
    """The whole per-GOP hot path from raw points (host or device) to bitstreams.
    Returns (EncodedGop, OptimState to seed the next GOP, per-epoch losses); with a `GopCoder` the first element is a
    future and the coding overlaps whatever the caller does next (normally the next GOP)."""
    frames = prepare_gop(points, scale_num, min_point_num, device)
    S = scale_num or frames[0].n_scales
    tr = GopTrainer(S, device, seed=seed, state=state, max_rows=max(f.tables.n_rows for f in frames), **(trainer_kwargs or {}))
    losses = tr.fit(frames, epochs)
    if coder is not None:
        return coder.submit(frames, tr.state.params, S), tr.state, losses
    enc = encode_gop(frames, tr.state.params, S, bitdepth, threads=threads)
    return enc, tr.state, losses


# ---- the reference's on-disk layout (encoder.py:13-18,84-146) ---------------------------------------------------
def write_gop(enc: EncodedGop, gop_dir: str):
    bins = os.path.join(gop_dir, "bins")
    os.makedirs(bins, exist_ok=True)
    with open(os.path.join(bins, "low_enc_bytes.bin"), "wb") as f:
        f.write(enc.low_enc_bytes)
    with open(os.path.join(bins, "model.bin"), "wb") as f:
        f.write(enc.model_bytes)
    with open(os.path.join(gop_dir, "side_info.json"), "w") as f:
        json.dump(enc.side_info, f, indent=4)
    for i, fb in enumerate(enc.frame_bytes):
        for s, b in enumerate(fb):
            with open(os.path.join(bins, f"frame{i:04d}_scale{s}.bin"), "wb") as f:
                f.write(b)


def read_gop(gop_dir: str, scale_num: int, n_frames: int) -> EncodedGop:
    bins = os.path.join(gop_dir, "bins")
    with open(os.path.join(gop_dir, "side_info.json")) as f:
        side = json.load(f)
    rd = lambda n: open(os.path.join(bins, n), "rb").read()
    frame_bytes = []
    for i in range(n_frames):
        fb, s = [], 0
        while os.path.exists(os.path.join(bins, f"frame{i:04d}_scale{s}.bin")):
            fb.append(rd(f"frame{i:04d}_scale{s}.bin"))
            s += 1
        frame_bytes.append(fb)
    return EncodedGop(scale_num, side, rd("model.bin"), 0.0, rd("low_enc_bytes.bin"), frame_bytes, [0] * n_frames)
