"""CPU stand-in for the MinkowskiEngine 0.5.4 symbols the reference touches.  TEST INFRASTRUCTURE ONLY.

Lets the reference's own `models/*.py` run on CPU in the build container so that
`oracle/gen_golden.py` can record fixtures of the reference control flow.  The arithmetic
is `oracle.linr_oracle.conv27` (ME itself is absent: conv numerics stay "parity unpinned").
Symbol list: SURVEY.md section 8(b).
"""
import math

import numpy as np
import torch
from torch import nn

from oracle import linr_oracle as O


class _Manager:
    def __init__(self):
        self.sets = {}      # key -> coords [N,4] int32
        self.nbr = {}       # key -> nbr27 table (int64 tensor)
        self._n = 0

    def insert(self, coords):
        coords = coords.to(torch.int32).contiguous()
        for k, c in self.sets.items():
            if c.shape == coords.shape and bool((c == coords).all()):
                return k
        k = ("k", self._n)
        self._n += 1
        self.sets[k] = coords
        return k

    def table(self, key):
        if key not in self.nbr:
            c = self.sets[key][:, 1:].numpy()
            # ME hashes coordinates: a lookup does not need sorted input
            order = np.argsort(O.pack_keys(c), kind="stable")
            srt = c[order]
            t = np.stack([O.lookup_rows(srt, c.astype(np.int64) + o) for o in O.OFFSETS27], axis=1)
            t = np.where(t >= 0, order[np.maximum(t, 0)], -1)
            self.nbr[key] = torch.from_numpy(t.astype(np.int64))
        return self.nbr[key]

    def cross_table(self, in_key, out_key):
        cin = self.sets[in_key][:, 1:].numpy()
        cout = self.sets[out_key][:, 1:].numpy().astype(np.int64)
        order = np.argsort(O.pack_keys(cin), kind="stable")
        srt = cin[order]
        t = np.stack([O.lookup_rows(srt, cout + o) for o in O.OFFSETS27], axis=1)
        t = np.where(t >= 0, order[np.maximum(t, 0)], -1)
        return torch.from_numpy(t.astype(np.int64))


class SparseTensor:
    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_map_key=None,
                 coordinate_manager=None, device=None):
        self.F = features
        if coordinate_manager is None:
            coordinate_manager = _Manager()
        self.coordinate_manager = coordinate_manager
        if coordinate_map_key is None:
            assert coordinates is not None
            coordinate_map_key = coordinate_manager.insert(coordinates)
        self.coordinate_map_key = coordinate_map_key
        if isinstance(tensor_stride, int):
            tensor_stride = [tensor_stride] * 3
        self.tensor_stride = list(tensor_stride)
        assert self.C.shape[0] == features.shape[0]

    @property
    def C(self):
        return self.coordinate_manager.sets[self.coordinate_map_key]

    @property
    def D(self):
        return 3

    @property
    def device(self):
        return self.F.device

    def _like(self, feats):
        return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager, tensor_stride=self.tensor_stride)

    def __add__(self, other):
        assert other.coordinate_manager is self.coordinate_manager
        if other.coordinate_map_key == self.coordinate_map_key:
            return self._like(self.F + other.F)
        # union map: self's rows first, then rows only in other  [UPSTREAM order unverified]
        a, b = self.C, other.C
        ka = O.pack_keys(a[:, 1:].numpy())
        kb = O.pack_keys(b[:, 1:].numpy())
        pos = {int(k): i for i, k in enumerate(ka)}
        extra = [i for i, k in enumerate(kb) if int(k) not in pos]
        coords = torch.cat([a, b[extra]], dim=0) if extra else a
        feats = torch.zeros(coords.shape[0], self.F.shape[1])
        feats[: a.shape[0]] += self.F
        rows = []
        nxt = a.shape[0]
        for i, k in enumerate(kb):
            if int(k) in pos:
                rows.append(pos[int(k)])
            else:
                rows.append(nxt)
                nxt += 1
        feats.index_add_(0, torch.tensor(rows, dtype=torch.long), other.F)
        return SparseTensor(feats, coordinates=coords, coordinate_manager=self.coordinate_manager,
                            tensor_stride=self.tensor_stride)

    def __iadd__(self, other):
        r = self + other
        self.F = r.F
        return self


class MinkowskiConvolution(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert stride == 1 and dilation == 1 and dimension == 3 and kernel_size in (1, 3)
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

    def forward(self, x, coordinates=None):
        m = x.coordinate_manager
        if coordinates is None:
            out_key = x.coordinate_map_key
        elif isinstance(coordinates, SparseTensor):
            out_key = coordinates.coordinate_map_key
        else:
            out_key = m.insert(coordinates)
        if self.kernel_size == 1:
            assert out_key == x.coordinate_map_key
            f = O.conv1(x.F, self.kernel, self.bias) if self.bias is not None else x.F @ self.kernel
        else:
            t = m.table(out_key) if out_key == x.coordinate_map_key else m.cross_table(x.coordinate_map_key, out_key)
            f = O.conv27(x.F, t, self.kernel, self.bias)
        return SparseTensor(f, coordinate_map_key=out_key, coordinate_manager=m, tensor_stride=x.tensor_stride)


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return x._like(torch.relu(x.F))


class MinkowskiPruning(nn.Module):
    def forward(self, x, mask):
        return SparseTensor(x.F[mask], coordinates=x.C[mask], coordinate_manager=x.coordinate_manager,
                            tensor_stride=x.tensor_stride)


def cat(*tensors):
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tensors[0]
    k = tensors[0].coordinate_map_key
    assert all(t.coordinate_map_key == k for t in tensors)
    return tensors[0]._like(torch.cat([t.F for t in tensors], dim=1))


class utils:  # noqa: N801  (ME.utils.sparse_collate, function_utils.py:15)
    @staticmethod
    def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
        cs, fs = [], []
        for b, (c, f) in enumerate(zip(coords, feats)):
            cs.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int32), c.to(torch.int32)], dim=1))
            fs.append(f)
        return torch.cat(cs, 0), torch.cat(fs, 0)
