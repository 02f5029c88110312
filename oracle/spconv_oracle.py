"""CPU restatement of the spconv semantics the reference relies on (TEST INFRASTRUCTURE, see oracle/__init__.py).

PARITY UNPINNED: the arithmetic lives in the third-party `spconv` package (`pip install spconv-cu113`, no version
pin — /root/reference/README.md:61-63), which is neither vendored nor installable offline, and the reference has
no test or golden vector for it.  This file restates spconv 2.x's published behaviour,
anchored on the reference's own call sites:

  * SparseConvTensor construction ............ ponder/models/sparse_unet/spconv_unet_v1m1_base.py:247-256
  * SubMConv3d (k3 / k5 stem / k1 proj) ...... :41-66, :111-119, :219-225
  * SparseConv3d(kernel 2, stride 2) ......... :135-142
  * SparseInverseConv3d(kernel 2, indice_key)  :171-177
  * the U-Net wiring / BN / ReLU / skip cat .. :70-83, :86-278

Rulebooks are integer numpy; convolutions are gather -> matmul -> index_add in torch (fp32 or fp64) so that
autograd supplies dgrad / wgrad for the backward parity tests.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------
# rulebooks (integer, bit-exact targets)
# ----------------------------------------------------------------------------------------------
def linear_key(coords: np.ndarray, shape) -> np.ndarray:
    c = coords.astype(np.int64)
    return ((c[:, 0] * int(shape[0]) + c[:, 1]) * int(shape[1]) + c[:, 2]) * int(shape[2]) + c[:, 3]


def subm_rulebook(coords: np.ndarray, shape, ksize: int) -> np.ndarray:
    """nbr[k][j] = smallest row i with coord[i] == coord[j] + (k - ksize//2), else -1;
    k = (k0*ksize + k1)*ksize + k2 over the three coordinate columns in stored order (Appendix B)."""
    n = coords.shape[0]
    K = ksize ** 3
    nbr = np.full((K, n), -1, dtype=np.int32)
    if n == 0:
        return nbr
    keys = linear_key(coords, shape)
    order = np.argsort(keys, kind="stable")  # stable: first occurrence = smallest row index
    skeys = keys[order]
    r = ksize // 2
    c = coords.astype(np.int64)
    k = 0
    for k0 in range(ksize):
        for k1 in range(ksize):
            for k2 in range(ksize):
                q = c.copy()
                q[:, 1] += k0 - r
                q[:, 2] += k1 - r
                q[:, 3] += k2 - r
                ok = ((q[:, 1] >= 0) & (q[:, 1] < shape[0]) & (q[:, 2] >= 0) & (q[:, 2] < shape[1])
                      & (q[:, 3] >= 0) & (q[:, 3] < shape[2]))
                qk = linear_key(q, shape)
                pos = np.searchsorted(skeys, qk, side="left")
                pos_c = np.minimum(pos, n - 1)
                hit = ok & (pos < n) & (skeys[pos_c] == qk)
                nbr[k, hit] = order[pos_c[hit]].astype(np.int32)
                k += 1
    return nbr


def down_rulebook(coords: np.ndarray, shape) -> Tuple[np.ndarray, np.ndarray, np.ndarray, List[int]]:
    """SparseConv3d(kernel 2, stride 2, padding 0): out coord = in >> 1, kernel offset = in & 1 per axis.

    Returns (out_coords [M,4] in CANONICAL order = ascending linearised (b,c0,c1,c2) key, in2out [N] (-1 when the
    2x2x2 window falls outside the unpadded input), koff [N], out_shape)."""
    out_shape = [(int(s) - 2) // 2 + 1 for s in shape]
    c = coords.astype(np.int64)
    q = c.copy()
    q[:, 1:] >>= 1
    koff = (((c[:, 1] & 1) * 2 + (c[:, 2] & 1)) * 2 + (c[:, 3] & 1)).astype(np.int32)
    ok = (q[:, 1] < out_shape[0]) & (q[:, 2] < out_shape[1]) & (q[:, 3] < out_shape[2])
    keys = linear_key(q, out_shape)
    uniq = np.unique(keys[ok])
    in2out = np.full(c.shape[0], -1, dtype=np.int32)
    in2out[ok] = np.searchsorted(uniq, keys[ok]).astype(np.int32)
    out = np.zeros((uniq.shape[0], 4), dtype=np.int32)
    rem = uniq.copy()
    out[:, 3] = rem % out_shape[2]; rem //= out_shape[2]
    out[:, 2] = rem % out_shape[1]; rem //= out_shape[1]
    out[:, 1] = rem % out_shape[0]; rem //= out_shape[0]
    out[:, 0] = rem
    return out, in2out, koff, out_shape


def down_maps(in2out: np.ndarray, koff: np.ndarray, n_out: int) -> Tuple[np.ndarray, np.ndarray]:
    n = in2out.shape[0]
    nbr_down = np.full((8, n_out), -1, dtype=np.int32)
    nbr_up = np.full((8, n), -1, dtype=np.int32)
    for i in range(n - 1, -1, -1):  # reverse so that the smallest fine row wins on duplicates
        if in2out[i] >= 0:
            nbr_down[koff[i], in2out[i]] = i
    rows = np.arange(n)
    nbr_up[koff, rows] = in2out
    return nbr_down, nbr_up


def canonical_down(out_coords: np.ndarray, in2out: np.ndarray, out_shape) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Bring an implementation-ordered strided rulebook into the canonical (ascending key) order.
    Returns (sorted out_coords, remapped in2out, perm) with sorted = out_coords[perm]."""
    keys = linear_key(out_coords, out_shape)
    perm = np.argsort(keys, kind="stable")
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.shape[0])
    remapped = np.where(in2out >= 0, inv[np.maximum(in2out, 0)], -1).astype(np.int32)
    return out_coords[perm], remapped, perm


# ----------------------------------------------------------------------------------------------
# arithmetic
# ----------------------------------------------------------------------------------------------
def sparse_conv(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], nbr: np.ndarray,
                n_out: int) -> torch.Tensor:
    """y[j] = bias + sum_k W[:,k,:] x[nbr[k][j]];  weight is [Cout, k0, k1, k2, Cin] (spconv KRSC)."""
    cout, cin = weight.shape[0], weight.shape[-1]
    w3 = weight.reshape(cout, -1, cin)
    y = torch.zeros(n_out, cout, dtype=x.dtype)
    for k in range(w3.shape[1]):
        rows = np.nonzero(nbr[k] >= 0)[0]
        if rows.size == 0:
            continue
        src = torch.from_numpy(nbr[k][rows].astype(np.int64))
        dst = torch.from_numpy(rows.astype(np.int64))
        y = y.index_add(0, dst, x.index_select(0, src) @ w3[:, k, :].t())
    if bias is not None:
        y = y + bias
    return y


class OracleSparseTensor:
    def __init__(self, features, indices: np.ndarray, spatial_shape, batch_size, rulebooks=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = batch_size
        self.rulebooks: Dict[str, object] = rulebooks if rulebooks is not None else {}

    def replace_feature(self, f):
        return OracleSparseTensor(f, self.indices, self.spatial_shape, self.batch_size, self.rulebooks)


def subm_conv(x: OracleSparseTensor, weight, bias, ksize: int, key: Optional[str]) -> OracleSparseTensor:
    if ksize == 1:
        y = x.features @ weight.reshape(weight.shape[0], -1).t()
        if bias is not None:
            y = y + bias
        return x.replace_feature(y)
    rb = x.rulebooks.get(key) if key is not None else None
    if rb is None:
        rb = subm_rulebook(x.indices, x.spatial_shape, ksize)
        if key is not None:
            x.rulebooks[key] = rb
    return x.replace_feature(sparse_conv(x.features, weight, bias, rb, x.indices.shape[0]))


def down_conv(x: OracleSparseTensor, weight, bias, key: str) -> OracleSparseTensor:
    out_coords, in2out, koff, out_shape = down_rulebook(x.indices, x.spatial_shape)
    nbr_down, nbr_up = down_maps(in2out, koff, out_coords.shape[0])
    x.rulebooks[key] = (x.indices, x.spatial_shape, nbr_up)
    y = sparse_conv(x.features, weight, bias, nbr_down, out_coords.shape[0])
    return OracleSparseTensor(y, out_coords, out_shape, x.batch_size, x.rulebooks)


def inverse_conv(x: OracleSparseTensor, weight, bias, key: str) -> OracleSparseTensor:
    fine_indices, fine_shape, nbr_up = x.rulebooks[key]
    y = sparse_conv(x.features, weight, bias, nbr_up, fine_indices.shape[0])
    return OracleSparseTensor(y, fine_indices, fine_shape, x.batch_size, x.rulebooks)


# ----------------------------------------------------------------------------------------------
# SpUNet-v1m1 wiring (spconv_unet_v1m1_base.py:86-278), functional over a reference-keyed state_dict
# ----------------------------------------------------------------------------------------------
def _bn_train(x, w, b, eps=1e-3):
    mean = x.mean(0)
    var = x.var(0, unbiased=False)
    return (x - mean) / torch.sqrt(var + eps) * w + b


def _basic_block(x: OracleSparseTensor, sd, prefix: str, key: str) -> OracleSparseTensor:
    """BasicBlock.forward, spconv_unet_v1m1_base.py:70-83."""
    residual = x
    out = subm_conv(x, sd[prefix + "conv1.weight"], None, 3, key)
    out = out.replace_feature(torch.relu(_bn_train(out.features, sd[prefix + "bn1.weight"], sd[prefix + "bn1.bias"])))
    out = subm_conv(out, sd[prefix + "conv2.weight"], None, 3, key)
    out = out.replace_feature(_bn_train(out.features, sd[prefix + "bn2.weight"], sd[prefix + "bn2.bias"]))
    if (prefix + "proj.0.weight") in sd:
        r = subm_conv(residual, sd[prefix + "proj.0.weight"], None, 1, None)
        res = _bn_train(r.features, sd[prefix + "proj.1.weight"], sd[prefix + "proj.1.bias"])
    else:
        res = residual.features
    return out.replace_feature(torch.relu(out.features + res))


def offset2batch(offset: np.ndarray) -> np.ndarray:
    """ponder/models/utils.py:11-26."""
    counts = np.diff(np.concatenate([[0], np.asarray(offset, dtype=np.int64)]))
    return np.repeat(np.arange(len(counts)), counts)


def spunet_forward(sd: Dict[str, torch.Tensor], grid_coord: np.ndarray, feat: torch.Tensor, offset: np.ndarray,
                   layers=(2, 3, 4, 6, 2, 2, 2, 2), return_levels: bool = False):
    """SpUNetBase.forward (num_classes=0, cls_mode=False), spconv_unet_v1m1_base.py:242-278, BN in train mode.

    Strided levels use the canonical (ascending-key) row order; the final features are returned in input row order,
    which is order-independent."""
    batch = offset2batch(offset)
    shape = (grid_coord.max(0) + 96).tolist()
    indices = np.concatenate([batch[:, None], grid_coord], 1).astype(np.int32)
    x = OracleSparseTensor(feat, indices, shape, int(batch[-1]) + 1)
    num_stages = len(layers) // 2
    x = subm_conv(x, sd["conv_input.0.weight"], None, 5, "stem")
    x = x.replace_feature(torch.relu(_bn_train(x.features, sd["conv_input.1.weight"], sd["conv_input.1.bias"])))
    skips = [x]
    for s in range(num_stages):
        x = down_conv(x, sd[f"down.{s}.0.weight"], None, f"spconv{s + 1}")
        x = x.replace_feature(torch.relu(_bn_train(x.features, sd[f"down.{s}.1.weight"], sd[f"down.{s}.1.bias"])))
        for i in range(layers[s]):
            x = _basic_block(x, sd, f"enc.{s}.block{i}.", f"subm{s + 1}")
        skips.append(x)
    levels = [t.indices for t in skips]
    x = skips.pop(-1)
    for s in reversed(range(num_stages)):
        x = inverse_conv(x, sd[f"up.{s}.0.weight"], None, f"spconv{s + 1}")
        x = x.replace_feature(torch.relu(_bn_train(x.features, sd[f"up.{s}.1.weight"], sd[f"up.{s}.1.bias"])))
        skip = skips.pop(-1)
        x = x.replace_feature(torch.cat((x.features, skip.features), dim=1))
        for i in range(layers[len(layers) - s - 1]):
            x = _basic_block(x, sd, f"dec.{s}.block{i}.", f"subm{s}")
    if return_levels:
        return x.features, levels
    return x.features
