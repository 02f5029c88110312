"""Seeded synthetic inputs of the BASELINE.json shapes: surface-like voxel clouds and ray batches.

No dataset or checkpoint is available offline, so every benchmark and parity test runs on these generators.
All randomness comes from `numpy.random.default_rng(seed)`; the same seed gives the same cloud on every box.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def _unique_rows(vox: np.ndarray, nrm: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    key = (vox[:, 0].astype(np.int64) << 42) | (vox[:, 1].astype(np.int64) << 21) | vox[:, 2].astype(np.int64)
    _, first = np.unique(key, return_index=True)
    first.sort()
    return vox[first], nrm[first]


def _plane(axis: int, level: int, lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
    axes = [a for a in range(3) if a != axis]
    u = np.arange(lo[axes[0]], hi[axes[0]] + 1)
    v = np.arange(lo[axes[1]], hi[axes[1]] + 1)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    out = np.empty((uu.size, 3), dtype=np.int64)
    out[:, axis] = level
    out[:, axes[0]] = uu.ravel()
    out[:, axes[1]] = vv.ravel()
    return out


def indoor_cloud(n: int, seed: int, grid_size: float = 0.02) -> Dict[str, np.ndarray]:
    """ScanNet/Structured3D-shaped scene: six room faces (8x6x3 m aspect) plus 20-40 axis-aligned rectangular
    patches, voxelised; the extent is scaled until exactly `n` distinct voxels remain."""
    rng = np.random.default_rng(seed)
    base = np.array([400.0, 300.0, 150.0])  # 8 x 6 x 3 m at 2 cm
    n_patch = int(rng.integers(20, 41))
    patch_spec = [(int(rng.integers(0, 3)), rng.random(), rng.random(2), 0.05 + 0.25 * rng.random(2))
                  for _ in range(n_patch)]
    scale = np.sqrt(n / 550000.0)
    for _ in range(30):
        dims = np.maximum(np.round(base * scale).astype(np.int64), 4)
        lo, hi = np.zeros(3, dtype=np.int64), dims - 1
        parts, norms = [], []
        for axis in range(3):
            for level in (0, int(hi[axis])):
                p = _plane(axis, level, lo, hi)
                parts.append(p)
                nv = np.zeros(3); nv[axis] = 1.0 if level == 0 else -1.0
                norms.append(np.broadcast_to(nv, p.shape))
        for axis, lev, org, ext in patch_spec:
            axes = [a for a in range(3) if a != axis]
            plo, phi = lo.copy(), hi.copy()
            for t, a in enumerate(axes):
                size = max(int(ext[t] * dims[a]), 2)
                start = int(org[t] * max(dims[a] - size, 1))
                plo[a], phi[a] = start, min(start + size - 1, hi[a])
            p = _plane(axis, int(lev * (dims[axis] - 1)), plo, phi)
            parts.append(p)
            nv = np.zeros(3); nv[axis] = 1.0
            norms.append(np.broadcast_to(nv, p.shape))
        vox, nrm = _unique_rows(np.concatenate(parts), np.concatenate(norms))
        if vox.shape[0] >= n:
            break
        scale *= 1.08
    keep = rng.permutation(vox.shape[0])[:n]
    vox, nrm = vox[keep], nrm[keep]
    vox = vox - vox.min(0)
    color = rng.random((n, 3))
    return {
        "grid_coord": vox.astype(np.int64),
        "coord": (vox * grid_size).astype(np.float32),
        "feat": np.concatenate([color, nrm], 1).astype(np.float32),
        "offset": np.array([n], dtype=np.int64),
    }


def outdoor_cloud(n: int, seed: int, voxel: float = 0.1) -> Dict[str, np.ndarray]:
    """nuScenes-shaped sweep: 32 lidar rings on a ground plane plus 30 boxes inside a 108 x 108 x 8 m extent."""
    rng = np.random.default_rng(seed)
    pts = []
    m = max(n // 16, 1024)
    for ring in range(32):
        r = 3.0 + 50.0 * (ring + 1) / 32.0
        th = rng.random(m) * 2 * np.pi
        pts.append(np.stack([54 + r * np.cos(th), 54 + r * np.sin(th), 1.0 + 0.05 * rng.standard_normal(m)], 1))
    for _ in range(30):
        c = np.array([rng.uniform(10, 98), rng.uniform(10, 98), 1.0])
        s = np.array([rng.uniform(1.5, 5), rng.uniform(1.5, 5), rng.uniform(1, 3)])
        q = rng.random((m // 4, 3))
        face = rng.integers(0, 3, m // 4)
        q[np.arange(m // 4), face] = rng.integers(0, 2, m // 4)
        pts.append(c + (q - [0.5, 0.5, 0.0]) * s)
    p = np.concatenate(pts)
    p = p[(p[:, 0] >= 0) & (p[:, 0] < 108) & (p[:, 1] >= 0) & (p[:, 1] < 108) & (p[:, 2] >= 0) & (p[:, 2] < 8)]
    vox = np.floor(p / voxel).astype(np.int64)
    vox, _ = _unique_rows(vox, np.zeros((vox.shape[0], 3)))
    if vox.shape[0] < n:
        extra = np.stack([rng.integers(0, 1080, 2 * n), rng.integers(0, 1080, 2 * n), rng.integers(8, 14, 2 * n)], 1)
        vox, _ = _unique_rows(np.concatenate([vox, extra]), np.zeros((vox.shape[0] + 2 * n, 3)))
    keep = rng.permutation(vox.shape[0])[:n]
    vox = vox[keep]
    coord = ((vox + 0.5) * voxel).astype(np.float32)
    strength = rng.random((n, 1)).astype(np.float32)
    return {
        "grid_coord": (vox - vox.min(0)).astype(np.int64),
        "coord": coord,
        "feat": np.concatenate([coord, strength], 1).astype(np.float32),
        "offset": np.array([n], dtype=np.int64),
    }


def ray_batch(r: int, seed: int, bbox=(-0.55, -0.55, -0.55, 0.55, 0.55, 0.55)) -> Dict[str, np.ndarray]:
    """Origins uniform inside the middle half of the AABB, directions uniform on the sphere, U(0.1,1) depth and
    U(0,1) colour targets."""
    rng = np.random.default_rng(seed)
    lo, hi = np.array(bbox[:3]), np.array(bbox[3:])
    mid, half = (lo + hi) / 2, (hi - lo) / 4
    o = mid + (rng.random((r, 3)) * 2 - 1) * half
    d = rng.standard_normal((r, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return {
        "rays_o": o.astype(np.float32), "rays_d": d.astype(np.float32),
        "depth": rng.uniform(0.1, 1.0, (r, 1)).astype(np.float32),
        "rgb": rng.random((r, 3)).astype(np.float32),
    }
