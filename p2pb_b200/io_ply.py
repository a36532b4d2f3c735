"""Minimal PLY / XYZ point-cloud I/O (open3d, which the reference uses, is not a dependency)."""
from __future__ import annotations

import numpy as np

_DT = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "char": "i1",
       "int": "i4", "int32": "i4", "uint": "u4", "short": "i2", "ushort": "u2"}


def read_ply(path: str):
    """-> (points [N,3] float64, colors [N,3] float64 in [0,1] or None). ascii or binary_little_endian vertex element."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply", "not a PLY file"
        fmt, props, n, in_vertex = None, [], 0, False
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element"):
                in_vertex = line.split()[1] == "vertex"
                if in_vertex:
                    n = int(line.split()[2])
            elif line.startswith("property") and in_vertex:
                t = line.split()
                props.append((t[-1], _DT[t[1]]))
            elif line == "end_header":
                break
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=n, ndmin=2)
            cols = {name: data[:, i] for i, (name, _) in enumerate(props)}
        else:
            dt = np.dtype([(name, ("<" if fmt.endswith("little_endian") else ">") + t) for name, t in props])
            arr = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
            cols = {name: arr[name] for name, _ in props}
    pts = np.stack([cols["x"], cols["y"], cols["z"]], 1).astype(np.float64)
    colors = None
    if all(k in cols for k in ("red", "green", "blue")):
        c = np.stack([cols["red"], cols["green"], cols["blue"]], 1).astype(np.float64)
        colors = c / 255.0 if c.max() > 1.0 else c
    return pts, colors


def write_ply(path: str, points: np.ndarray, colors: np.ndarray = None) -> None:
    """binary_little_endian PLY with double xyz (+ uchar rgb), what open3d writes for a PointCloud."""
    n = points.shape[0]
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {n}", "property double x", "property double y", "property double z"]
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if colors is not None:
        hdr += ["property uchar red", "property uchar green", "property uchar blue"]
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    hdr.append("end_header")
    arr = np.empty(n, dtype=np.dtype(fields))
    arr["x"], arr["y"], arr["z"] = points[:, 0], points[:, 1], points[:, 2]
    if colors is not None:
        c = np.clip(np.round(colors * 255.0), 0, 255).astype(np.uint8)
        arr["red"], arr["green"], arr["blue"] = c[:, 0], c[:, 1], c[:, 2]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        f.write(arr.tobytes())


def write_array_to_xyz(path: str, array: np.ndarray) -> None:
    """Same text format as the reference (utils/utils.py:5-10): '%8f' per value, no trailing newline."""
    fmt = "\n".join([" ".join(["%8f"] * array.shape[1])] * array.shape[0])
    with open(path, "w") as f:
        f.write(fmt % tuple(array.ravel()))
