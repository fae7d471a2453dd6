"""The MinkowskiEngine 0.5.4 symbols LINR-PCGC's hot path touches, on the sm_100a kernels.

    SparseTensor(features, coordinates=None, tensor_stride=1, coordinate_map_key=None, coordinate_manager=None,
                 device=None)      .F .C .D .tensor_stride .coordinate_manager .coordinate_map_key .device, + / +=
    MinkowskiConvolution(in_channels, out_channels, kernel_size, stride, dilation, bias, dimension)
                                   conv(x) and conv(x, coordinates)   (models/upsample.py:17-23,90-97; models/resnet.py:15-51)
    MinkowskiReLU(inplace)         models/upsample.py:92; models/resnet.py:53
    MinkowskiPruning()(x, mask)    models/upsample.py:116
    cat(a, b) / cat([a, b])        models/resnet.py:58
    utils.sparse_collate           models/function_utils.py:15

Semantics restated from ME's published behaviour (absent from /root/reference): cross-correlation,
offset k = (dx+1) + 3(dy+1) + 9(dz+1), kernels [27,Cin,Cout] (2-D [Cin,Cout] for kernel_size 1), bias [1,Cout],
init U(+-1/sqrt(Cin*K)), unique input coordinates keep their row order.  What differs from ME on purpose: the
27-neighbour kernel map of a coordinate set is built ONCE per coordinate tensor (cached by storage, not per
SparseTensor / manager), and the backward pass has no floating-point atomics (bitwise reproducible).

Only what the live network uses is implemented: dimension 3, stride 1, dilation 1, kernel_size 1 or 3, one batch.
Anything else raises NotImplementedError.  CUDA tensors only.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Optional

import torch
from torch import nn

from ... import net as _net
from ...frame import RowTables, build_tables

__version__ = "0.5.4+linr_b200"


def _check_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("linr_b200 MinkowskiEngine shim: tensors must live on a CUDA device (there is no CPU path)")


# ---------------------------------------------------------------------------------------------- kernel-map cache
class _TableCache:
    """coordinate storage -> RowTables.  The reference wraps the same coordinate tensors into fresh SparseTensors
    ~16x per scale per iteration (models/function_utils.py:13-18,58-69); the kernel map only depends on the set."""

    def __init__(self, capacity: int = 512):
        self.cap = capacity
        self.d: "OrderedDict[tuple, tuple]" = OrderedDict()

    @staticmethod
    def ident(t: torch.Tensor) -> tuple:
        src = getattr(t, "_linr_src", None)        # set by utils.sparse_collate: identity of the un-batched tensor
        return src if src is not None else (t.data_ptr(), int(t.shape[0]), int(t.shape[1]), t._version)

    def get(self, coords: torch.Tensor) -> RowTables:
        key = self.ident(coords)
        hit = self.d.get(key)
        if hit is not None:
            self.d.move_to_end(key)
            return hit[0]
        xyz = coords[:, -3:].to(torch.int32).contiguous()
        _check_coordinate_set(coords, xyz)
        scale = torch.zeros(int(xyz.shape[0]), dtype=torch.uint8, device=xyz.device)
        tab = build_tables(xyz, scale)
        self.d[key] = (tab, coords)                 # keeps the key tensor (and, through `_linr_keep`, the un-batched
                                                    # source) alive: their storage pointers cannot be recycled
        if len(self.d) > self.cap:
            self.d.popitem(last=False)
        return tab


def _check_coordinate_set(coords: torch.Tensor, xyz: torch.Tensor) -> None:
    """The compact kernel map (9 anchors + 27 bits per row) and the hash insert assume ONE batch of x-major sorted,
    unique rows -- the reference's invariant for every set it convolves (datautils/custom_dataset.py:308,
    models/sort_functions.py:95-103).  Checked once per coordinate set (the tables are cached); anything else would
    gather wrong rows silently, so it raises."""
    n = int(xyz.shape[0])
    if n == 0:
        return
    if coords.shape[1] == 4 and bool((coords[:, 0] != coords[0, 0]).any()):
        raise NotImplementedError("linr_b200 MinkowskiEngine shim: one batch index per coordinate set (LINR-PCGC collates single frames)")
    if int(xyz.min()) < 0 or int(xyz.max()) >= (1 << 20):
        raise ValueError("linr_b200 MinkowskiEngine shim: coordinates must be in [0, 2^20)")
    if n > 1:
        k = (xyz[:, 0].to(torch.int64) << 40) | (xyz[:, 1].to(torch.int64) << 20) | xyz[:, 2].to(torch.int64)
        d = k[1:] - k[:-1]
        if bool((d <= 0).any()):
            what = "duplicate rows" if bool((d == 0).any()) and not bool((d < 0).any()) else "rows that are not x-major sorted"
            raise ValueError(f"linr_b200 MinkowskiEngine shim: coordinate set with {what}; sort it first "
                             "(models/sort_functions.py sort_sparse_tensor) -- the kernel map needs sorted unique rows")


_tables = _TableCache()


class CoordinateManager:
    """Coordinate sets of one SparseTensor family: key -> [N,1+3] int32 (batch column first)."""

    def __init__(self, D: int = 3):
        if D != 3:
            raise NotImplementedError("only 3-D coordinates")
        self.sets: dict = {}
        self._n = 0

    def insert(self, coords: torch.Tensor):
        if coords.dim() != 2 or coords.shape[1] != 4:
            raise ValueError("coordinates must be [N, 1+3] with the batch index in column 0")
        ident = _TableCache.ident(coords)
        for k, (c, cid) in self.sets.items():
            if c is coords or cid == ident:
                return k
        for k, (c, cid) in self.sets.items():       # same content under another tensor (one sync, rare)
            if c.shape == coords.shape and bool(torch.equal(c, coords)):
                return k
        k = ("key", self._n)
        self._n += 1
        c = coords if coords.dtype == torch.int32 else coords.to(torch.int32)
        if c is not coords and hasattr(coords, "_linr_src"):
            c._linr_src, c._linr_keep = coords._linr_src, coords._linr_keep
        self.sets[k] = (c, ident)
        return k

    def coords(self, key) -> torch.Tensor:
        return self.sets[key][0]

    def table(self, key) -> RowTables:
        return _tables.get(self.sets[key][0])


# ---------------------------------------------------------------------------------------------- SparseTensor
class SparseTensor:
    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_map_key=None,
                 coordinate_manager: Optional[CoordinateManager] = None, device=None, **unsupported):
        if unsupported:
            raise NotImplementedError(f"SparseTensor arguments not used by LINR-PCGC: {sorted(unsupported)}")
        _check_cuda(features)
        self.F = features
        if coordinate_manager is None:
            coordinate_manager = CoordinateManager()
        self.coordinate_manager = coordinate_manager
        if coordinate_map_key is None:
            if coordinates is None:
                raise ValueError("either coordinates or coordinate_map_key is required")
            coordinate_map_key = coordinate_manager.insert(coordinates.to(features.device))
        self.coordinate_map_key = coordinate_map_key
        if isinstance(tensor_stride, int):
            tensor_stride = [tensor_stride] * 3
        self.tensor_stride = list(tensor_stride)
        if self.C.shape[0] != features.shape[0]:
            raise ValueError(f"{features.shape[0]} feature rows for {self.C.shape[0]} coordinates")

    # -- attributes the reference reads
    @property
    def C(self) -> torch.Tensor:
        return self.coordinate_manager.coords(self.coordinate_map_key)

    @property
    def D(self) -> int:
        return 3

    @property
    def device(self):
        return self.F.device

    @property
    def shape(self):
        return self.F.shape

    def __len__(self):
        return int(self.F.shape[0])

    def _like(self, feats) -> "SparseTensor":
        return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key, coordinate_manager=self.coordinate_manager,
                            tensor_stride=self.tensor_stride)

    # -- x + y: same key -> elementwise; other key of the same manager -> union of the coordinate sets
    def __add__(self, other: "SparseTensor") -> "SparseTensor":
        if other.coordinate_manager is not self.coordinate_manager:
            raise ValueError("SparseTensors of different coordinate managers cannot be added")
        if other.coordinate_map_key == self.coordinate_map_key:
            return self._like(self.F + other.F)
        a, b = self.C, other.C
        # union map in x-major sorted order (torch.unique sorts rows lexicographically, batch column first): a union of
        # sorted sets stays a valid input of the 27-offset convolution; equal sets keep their row order
        both = torch.cat([a, b], dim=0)
        coords, inv = torch.unique(both, dim=0, return_inverse=True)
        feats = torch.zeros((coords.shape[0], self.F.shape[1]), dtype=self.F.dtype, device=self.F.device)
        feats.index_add_(0, inv[: a.shape[0]], self.F)
        feats.index_add_(0, inv[a.shape[0]:], other.F)
        return SparseTensor(feats, coordinates=coords, coordinate_manager=self.coordinate_manager,
                            tensor_stride=self.tensor_stride)

    def __iadd__(self, other: "SparseTensor") -> "SparseTensor":
        r = self + other
        self.F, self.coordinate_map_key = r.F, r.coordinate_map_key
        return self

    def __repr__(self):
        return f"SparseTensor(N={len(self)}, C={int(self.F.shape[1])}, stride={self.tensor_stride}, key={self.coordinate_map_key})"


# ---------------------------------------------------------------------------------------------- convolution
def _pad_channels(c: int) -> int:
    if c <= 4:
        return 4
    if c <= 8:
        return 8
    raise NotImplementedError(f"linr_b200 sparse conv supports up to 8 channels (got {c}); LINR-PCGC uses hidden_channel_conv=8")


class _Conv27(torch.autograd.Function):
    """y = sum_k x[row(C + delta_k)] @ W[k] (+ bias): forward, grad-input (mirrored-offset gather) and the deterministic
    weight gradient, each one C-ABI call (linr_spconv27_fwd / _bwd_in / _bwd_w)."""

    @staticmethod
    def forward(ctx, x, kernel, bias, tables):
        cin, cout = int(kernel.shape[1]), int(kernel.shape[2])
        pi, po = _pad_channels(cin), _pad_channels(cout)
        xp = x if pi == cin else torch.nn.functional.pad(x, (0, pi - cin))
        wp = kernel if (pi == cin and po == cout) else torch.nn.functional.pad(kernel, (0, po - cout, 0, pi - cin))
        bp = None
        if bias is not None:
            bp = bias.reshape(-1)
            bp = bp if po == cout else torch.nn.functional.pad(bp, (0, po - cout))
        y = _net.spconv27_fwd(xp.contiguous(), wp.contiguous(), bp.contiguous() if bp is not None else None, tables)
        ctx.save_for_backward(xp, wp)
        ctx.tables, ctx.dims, ctx.has_bias = tables, (cin, cout, pi, po), bias is not None
        return y if po == cout else y[:, :cout].contiguous()

    @staticmethod
    def backward(ctx, dy):
        xp, wp = ctx.saved_tensors
        cin, cout, pi, po = ctx.dims
        dyp = dy.contiguous() if po == cout else torch.nn.functional.pad(dy, (0, po - cout)).contiguous()
        dx = dk = db = None
        if ctx.needs_input_grad[0]:
            dx = _net.spconv27_bwd_in(dyp, wp, ctx.tables)
            dx = dx if pi == cin else dx[:, :cin].contiguous()
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            if pi == 4 and po == 8:      # the weight-gradient entry point has no 4 -> 8 variant: widen the input
                dW, dbv = _net.spconv27_bwd_w(torch.nn.functional.pad(xp, (0, 4)).contiguous(), dyp, ctx.tables)
            else:
                dW, dbv = _net.spconv27_bwd_w(xp, dyp, ctx.tables)
            dk = dW[:, :cin, :cout].contiguous()
            if ctx.has_bias:
                db = dbv[:cout].reshape(1, cout)
        return dx, dk, db, None


class MinkowskiConvolution(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        if dimension != 3 or stride != 1 or dilation != 1 or kernel_size not in (1, 3) or kernel_generator is not None \
                or expand_coordinates:
            raise NotImplementedError("linr_b200 shim: only dimension=3, stride=1, dilation=1, kernel_size in {1,3} "
                                      "(all that LINR-PCGC's live network uses, SURVEY.md 0)")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        vol = kernel_size ** 3
        shape = (vol, in_channels, out_channels) if vol > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        stdv = 1.0 / math.sqrt(in_channels * vol)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, x: SparseTensor, coordinates=None) -> SparseTensor:
        m = x.coordinate_manager
        if coordinates is None:
            out_key = x.coordinate_map_key
        elif isinstance(coordinates, SparseTensor):
            out_key = coordinates.coordinate_map_key
        else:
            out_key = m.insert(coordinates)
        if out_key != x.coordinate_map_key:
            raise NotImplementedError("output coordinates other than the input set (LINR-PCGC always passes the same "
                                      "set, models/upsample.py:191)")
        if self.kernel_size == 1:
            f = x.F @ self.kernel
            if self.bias is not None:
                f = f + self.bias
        else:
            f = _Conv27.apply(x.F, self.kernel, self.bias, m.table(out_key))
        return x._like(f)

    def extra_repr(self):
        return f"in={self.in_channels}, out={self.out_channels}, kernel_size={self.kernel_size}"


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace: bool = False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x: SparseTensor) -> SparseTensor:
        return x._like(torch.relu(x.F))


class MinkowskiPruning(nn.Module):
    def forward(self, x: SparseTensor, mask: torch.Tensor) -> SparseTensor:
        if mask.dtype != torch.bool or mask.shape[0] != len(x):
            raise ValueError("pruning mask must be a bool vector with one entry per row")
        return SparseTensor(x.F[mask], coordinates=x.C[mask], coordinate_manager=x.coordinate_manager,
                            tensor_stride=x.tensor_stride)


def cat(*tensors) -> SparseTensor:
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    k = tensors[0].coordinate_map_key
    if any(t.coordinate_map_key != k or t.coordinate_manager is not tensors[0].coordinate_manager for t in tensors):
        raise ValueError("ME.cat needs tensors on the same coordinate map")
    return tensors[0]._like(torch.cat([t.F for t in tensors], dim=1))


class utils:  # noqa: N801  (ME.utils.sparse_collate, models/function_utils.py:15)
    @staticmethod
    def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
        if labels is not None or len(coords) != 1:
            raise NotImplementedError("one batch, no labels (LINR-PCGC codes one frame at a time)")
        c, f = coords[0], feats[0]
        out = torch.cat([torch.zeros((c.shape[0], 1), dtype=torch.int32, device=c.device), c.to(torch.int32)], dim=1)
        # the un-batched tensor identifies the set: re-wrapping the same coordinates re-uses their kernel map
        out._linr_src = ("xyz", c.data_ptr(), int(c.shape[0]), c._version)
        out._linr_keep = c     # whoever holds `out` (SparseTensor, kernel-map cache) keeps the source storage alive
        return out, f
