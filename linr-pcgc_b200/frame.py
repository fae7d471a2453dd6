"""Per-frame octree preparation on the GPU: the data every later kernel reads.

Host-side mirror of `MyDataset.handle_data` (datautils/custom_dataset.py:259-355), `qscTensor`
(models/module_utils.py:155-224) and `octree_level` (models/module_utils.py:86-152) of the reference,
but built once per frame and kept resident in HBM as packed tables:

    coords  int32 [R,3]   parents of every scale, scale 0 first (R = sum_s N_s)
    scale   uint8 [R]     scale index of each row
    occ     uint8 [R]     8-bit child occupancy, bit i = octant 4dx+2dy+dz (models/module_utils.py:93)
    nbr7    uint8 [R]     self + 6 face neighbours present (offsets_ini order, main.py:24)
    anchor  int32 [9,ld], mask uint32 [R]   compact 27-neighbour kernel map (see include/linr_b200.h)

About 54 B per row instead of the reference's pickled float tensors; a 32-frame loot GOP is ~0.5 GB.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import Rows, check, ptr, stream_ptr


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


@dataclass
class RowTables:
    """Kernel map + per-row side data of one coordinate set (one scale, or all scales of a frame)."""
    coords: torch.Tensor
    scale: torch.Tensor
    nbr7: torch.Tensor
    anchor: torch.Tensor
    mask: torch.Tensor
    occ: Optional[torch.Tensor] = None
    nbr27: Optional[torch.Tensor] = None
    tile_rng: Optional[torch.Tensor] = None   # int32 [ceil(R/128),6]: neighbour row ranges per tile (linr_tile_ranges)
    pair_cnt: Optional[torch.Tensor] = None   # int32 [ceil(R/256),32] and
    pair_list: Optional[torch.Tensor] = None  # int32 [ceil(R/256),27*256]: existing (row, neighbour) pairs per tile and offset (layout: linr_b200.h)
    table: Optional[torch.Tensor] = None   # open-addressing hash (kept only on request)
    cap: int = 0
    _rows: Optional[Rows] = field(default=None, repr=False)

    @property
    def n_rows(self) -> int:
        return int(self.coords.shape[0])

    def rows(self) -> Rows:
        """ctypes `linr_rows` view (re-made if the occupancy tensor was swapped)."""
        r = Rows()
        r.n_rows = self.n_rows
        r.ld = int(self.anchor.shape[1])
        r.d_anchor, r.d_mask = ptr(self.anchor), ptr(self.mask)
        r.d_nbr7, r.d_scale = ptr(self.nbr7), ptr(self.scale)
        r.d_occ = ptr(self.occ) if self.occ is not None else None
        r.d_tile_rng = ptr(self.tile_rng) if self.tile_rng is not None else None
        r.d_pair_cnt = ptr(self.pair_cnt) if self.pair_cnt is not None else None
        r.d_pair_list = ptr(self.pair_list) if self.pair_list is not None else None
        self._rows = r
        return r


def build_tables(coords: torch.Tensor, scale: torch.Tensor, occ: Optional[torch.Tensor] = None,
                 dense: bool = False, keep_hash: bool = False, tile_ranges: bool = True) -> RowTables:
    """Hash the rows and build the 27-neighbour kernel map (replaces ME's coordinate manager)."""
    lib = _lib.load()
    assert coords.is_cuda and coords.dtype == torch.int32 and coords.dim() == 2 and coords.shape[1] == 3
    coords = coords.contiguous()
    n = int(coords.shape[0])
    dev = coords.device
    cap = 1 << max(4, int(2 * max(n, 1) - 1).bit_length())
    table = _ws(lib.linr_hash_bytes(cap), dev)
    s = stream_ptr()
    check(lib.linr_hash_build(ptr(coords), ptr(scale), n, ptr(table), cap, s), "linr_hash_build")
    ld = (max(n, 1) + 31) // 32 * 32
    anchor = torch.empty((9, ld), dtype=torch.int32, device=dev)
    mask = torch.zeros(ld, dtype=torch.int32, device=dev)   # padded to ld: bulk copies read whole 16-byte pieces
    nbr7 = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    nbr27 = torch.empty((n, 27), dtype=torch.int32, device=dev) if dense else None
    check(lib.linr_nbr_build(ptr(coords), ptr(scale), n, ptr(table), cap, ptr(nbr27) if dense else None, ptr(anchor), ld,
                             ptr(mask), ptr(nbr7), s), "linr_nbr_build")
    t = RowTables(coords=coords, scale=scale, nbr7=nbr7[:n] if n else nbr7[:0], anchor=anchor, mask=mask[:n] if n else mask[:0],
                  occ=occ, nbr27=nbr27, table=table if (keep_hash or dense) else None, cap=cap)
    if n and tile_ranges:
        rng = torch.empty(((n + 127) // 128, 6), dtype=torch.int32, device=dev)
        rows = t.rows()
        check(lib.linr_tile_ranges(C.byref(rows), ptr(rng), s), "linr_tile_ranges")
        t.tile_rng = rng
        if n < (1 << 24):
            nt = (n + 255) // 256
            t.pair_cnt = torch.empty((nt, 32), dtype=torch.int32, device=dev)
            t.pair_list = torch.empty((nt, 27 * 256), dtype=torch.int32, device=dev)
            rows = t.rows()
            check(lib.linr_pair_lists(C.byref(rows), ptr(t.pair_cnt), ptr(t.pair_list), s), "linr_pair_lists")
    return t


def hash_lookup(t: RowTables, query: torch.Tensor, query_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """QuickSearchCoord.search_coord_idx (models/module_utils.py:276-283): row of each query coordinate or -1."""
    lib = _lib.load()
    if t.table is None:
        raise _lib.LinrError("tables were built without keep_hash=True")
    q = query.to(torch.int32).contiguous()
    rows = torch.empty(int(q.shape[0]), dtype=torch.int32, device=q.device)
    check(lib.linr_hash_lookup(ptr(q), ptr(query_scale), int(q.shape[0]), ptr(t.table), t.cap, ptr(rows), stream_ptr()),
          "linr_hash_lookup")
    return rows


def sort_unique(xyz: torch.Tensor, bits: int) -> torch.Tensor:
    """torch.unique(dim=0) in x-major lexicographic order (datautils/custom_dataset.py:280)."""
    lib = _lib.load()
    n = int(xyz.shape[0])
    out = torch.empty_like(xyz)
    cnt = torch.zeros(1, dtype=torch.int64, device=xyz.device)
    ws = _ws(lib.linr_coord_ws_bytes(n), xyz.device)
    check(lib.linr_coord_sort_unique(ptr(xyz.contiguous()), n, bits, ptr(out), ptr(cnt), ptr(ws), ws.numel(), stream_ptr()),
          "linr_coord_sort_unique")
    return out[: int(cnt.item())]


def sort_rows(xyz: torch.Tensor, bits: int) -> torch.Tensor:
    """sort_by_coord_sum_c (models/sort_functions.py:17-30)."""
    lib = _lib.load()
    n = int(xyz.shape[0])
    out = torch.empty_like(xyz)
    ws = _ws(lib.linr_coord_ws_bytes(n), xyz.device)
    check(lib.linr_coord_sort(ptr(xyz.contiguous()), n, bits, ptr(out), ptr(ws), ws.numel(), stream_ptr()), "linr_coord_sort")
    return out


def octree_down(child: torch.Tensor, bits: int):
    """octree_level.forward (models/module_utils.py:97-115) -> (parent [N,3] int32, occ uint8 [N])."""
    lib = _lib.load()
    nc = int(child.shape[0])
    dev = child.device
    parent = torch.empty((nc, 3), dtype=torch.int32, device=dev)
    occ = torch.empty((nc + 3) // 4 * 4, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _ws(lib.linr_coord_ws_bytes(nc), dev)
    check(lib.linr_octree_down(ptr(child.contiguous()), nc, bits, ptr(parent), ptr(occ), ptr(cnt), ptr(ws), ws.numel(),
                               stream_ptr()), "linr_octree_down")
    n = int(cnt.item())
    return parent[:n], occ[:n]


def octree_up(parent: torch.Tensor, occ: torch.Tensor, bits: int) -> torch.Tensor:
    """octree_level.upper_layer (models/module_utils.py:117-127): sorted children of (parent, occupancy)."""
    lib = _lib.load()
    n = int(parent.shape[0])
    dev = parent.device
    off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    ws = _ws(lib.linr_coord_ws_bytes(8 * n + 8), dev)
    s = stream_ptr()
    parent, occ = parent.contiguous(), occ.contiguous()
    check(lib.linr_octree_up_count(ptr(occ), n, ptr(off), ptr(ws), ws.numel(), s), "linr_octree_up_count")
    n_child = int(off[n].item())
    child = torch.empty((n_child, 3), dtype=torch.int32, device=dev)
    check(lib.linr_octree_up_expand(ptr(parent), ptr(occ), ptr(off), n, n_child, bits, ptr(child), ptr(ws), ws.numel(), s),
          "linr_octree_up_expand")
    return child


@dataclass
class Frame:
    """One prepared frame: all scales concatenated (`tables`) + what the codec needs on the host."""
    tables: RowTables
    scale_off: List[int]          # row offset of each scale in `tables`, len S+1
    point_num: int
    coord_min: np.ndarray         # int32 [3]
    xyz: torch.Tensor             # [point_num,3] sorted unique, min-subtracted (the "ori"/ground truth)
    bits: int

    @property
    def n_scales(self) -> int:
        return len(self.scale_off) - 1

    def scale_coords(self, s: int) -> torch.Tensor:
        return self.tables.coords[self.scale_off[s]: self.scale_off[s + 1]]

    def scale_occ(self, s: int) -> torch.Tensor:
        return self.tables.occ[self.scale_off[s]: self.scale_off[s + 1]]

    def scale_nbr7(self, s: int) -> torch.Tensor:
        return self.tables.nbr7[self.scale_off[s]: self.scale_off[s + 1]]


def prepare_frame(points: torch.Tensor, scale_num: Optional[int] = None, min_point_num: int = 64,
                  bits: Optional[int] = None, dense: bool = False) -> Frame:
    """points: CUDA int32 [Np,3] (any order, duplicates allowed).  Scale loop of custom_dataset.py:289-344:
    stop after the scale whose parent count drops below `min_point_num`, or at `scale_num`."""
    lib = _lib.load()
    if not points.is_cuda:
        raise _lib.LinrError("prepare_frame takes a CUDA tensor (host points are copied by the caller)")
    pts = points[:, :3].to(torch.int32).contiguous()
    n = int(pts.shape[0])
    dev = pts.device
    sub = torch.empty_like(pts)
    mn = torch.empty(3, dtype=torch.int32, device=dev)
    check(lib.linr_coord_min_sub(ptr(pts), n, ptr(sub), ptr(mn), stream_ptr()), "linr_coord_min_sub")
    if bits is None:
        bits = max(1, int(sub.max().item()).bit_length())
    xyz = sort_unique(sub, bits)
    cur = xyz
    coords, occs = [], []
    cap = 100000 if scale_num is None else scale_num
    s = 0
    while s < cap:
        parent, occ = octree_down(cur, bits)
        coords.append(parent)
        occs.append(occ)
        if parent.shape[0] < min_point_num or s == cap - 1:
            break
        cur = parent
        s += 1
    off = [0]
    for c in coords:
        off.append(off[-1] + int(c.shape[0]))
    allc = torch.cat(coords, dim=0)
    allocc = torch.cat(occs, dim=0)
    scale = torch.cat([torch.full((int(c.shape[0]),), i, dtype=torch.uint8, device=dev) for i, c in enumerate(coords)])
    tables = build_tables(allc, scale, allocc, dense=dense)
    return Frame(tables=tables, scale_off=off, point_num=int(xyz.shape[0]), coord_min=mn.cpu().numpy(), xyz=xyz, bits=bits)
