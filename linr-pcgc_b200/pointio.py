"""Point-cloud file IO without open3d: PLY (ascii / binary_little_endian, x y z first) and NPY readers, the ASCII PLY
writer of the reference (datautils/custom_dataset.py:10-14,36-58) and a directory dataset in file-name order
(datautils/custom_dataset.py:230-257, minus the pickle cache: prepared frames stay resident in HBM instead)."""
from __future__ import annotations

import os
from typing import List

import numpy as np

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
              "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4",
              "float32": "f4", "float64": "f8"}


def read_ply(path: str, dtype="int32") -> np.ndarray:
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties on vertices are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        names = [p[0] for p in props]
        if names[:3] != ["x", "y", "z"]:
            raise ValueError(f"{path}: vertex properties must start with x y z (got {names[:3]})")
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=n, ndmin=2)[:, :3]
        elif fmt == "binary_little_endian":
            rec = np.dtype([(nm, "<" + t) for nm, t in props])
            raw = np.frombuffer(f.read(n * rec.itemsize), dtype=rec, count=n)
            data = np.stack([raw["x"], raw["y"], raw["z"]], axis=1)
        else:
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
    return np.ascontiguousarray(data).astype(dtype)


def write_ply_ascii(path: str, coords: np.ndarray, dtype="int32") -> None:
    """Same header and row format as the reference writer (datautils/custom_dataset.py:36-58)."""
    coords = np.asarray(coords).astype(dtype)
    with open(path, "w") as f:
        f.write(f"ply\nformat ascii 1.0\nelement vertex {coords.shape[0]}\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
        np.savetxt(f, coords, fmt="%d" if np.issubdtype(coords.dtype, np.integer) else "%g")


class PointDirectory:
    """Frames of a sequence = the files of `ori_dir` with extension `ori_type`, sorted by name."""

    def __init__(self, ori_dir: str, ori_type: str = "ply"):
        self.files: List[str] = sorted(os.path.join(ori_dir, f) for f in os.listdir(ori_dir) if f.endswith("." + ori_type))
        self.ori_type = ori_type
        if not self.files:
            raise FileNotFoundError(f"no *.{ori_type} files under {ori_dir}")

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i: int) -> np.ndarray:
        p = self.files[i]
        return np.load(p)[:, :3].astype(np.int32) if self.ori_type == "npy" else read_ply(p)
