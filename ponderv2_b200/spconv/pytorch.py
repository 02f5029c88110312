"""`spconv.pytorch`-compatible module surface (boundary B1) backed by libpv2_b200.

Mirrors exactly the part of spconv 2.x that the reference touches
(ponder/models/sparse_unet/spconv_unet_v1m1_base.py:11,21,38-66,111-119,134-145,170-180,219-225,249-256,270):
`SparseConvTensor`, `SparseModule`, `SparseSequential`, `SubMConv3d`, `SparseConv3d`, `SparseInverseConv3d`,
`Identity`.  Semantics follow SURVEY.md Appendix B:

* indices are `[N,4] int32 (batch, c0, c1, c2)`; weights are `[Cout, k0, k1, k2, Cin]`, bias `[Cout]`;
* SubMConv3d keeps the input index set and row order, is always centred (the `padding` argument is ignored,
  as the reference's `kernel_size=5, padding=1` stem requires);
* SparseConv3d supports the one configuration the reference uses (kernel 2, stride 2, padding 0); output rows are
  numbered by first contributing input row; SparseInverseConv3d reuses that rulebook through `indice_key`;
* rulebooks are cached per tensor under `indice_key` and shared by every conv with the same key.

All arithmetic runs in hand-written sm_100a kernels; there is no CPU path (CPU tensors raise).
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import List, Optional, Sequence, Union

import torch
from torch import nn

from .. import _lib

# development switch (A/B timing of the mask-sorted tile order); results never depend on it
USE_ROW_ORDER = os.environ.get("PV2_ROW_ORDER", "1") != "0"

__all__ = [
    "SparseConvTensor", "SparseModule", "SparseSequential", "SubMConv3d", "SparseConv3d",
    "SparseInverseConv3d", "Identity", "build_subm_rulebook", "build_down_rulebook",
]


# ----------------------------------------------------------------------------------------------
# rulebooks
# ----------------------------------------------------------------------------------------------
class TileMap:
    """A neighbour map prepared for the conv kernels: `nbr` [K, n] (tile order when `order` is set: nbr[k][pos] feeds
    output row order[pos]), `order` [n] int32 or None, `blk_active` [K, ceil(n/32)] uint8 or None."""

    __slots__ = ("nbr", "order", "blk_active")

    def __init__(self, nbr, order=None, blk_active=None):
        self.nbr, self.order, self.blk_active = nbr, order, blk_active


def build_tile_map(nbr: torch.Tensor) -> TileMap:
    """Rows sorted by neighbour-presence mask so that 128-row tiles touch few kernel offsets (pv2_rulebook_row_order),
    the map permuted into that order, and the per-32-row-block activity bytes of the weight-gradient kernel.
    Maps the tile-skipping kernels do not use (K > 32, K = 1, empty) stay in natural order."""
    kvol, n = nbr.shape
    if not USE_ROW_ORDER or kvol > 128 or kvol < 2 or n == 0:
        return TileMap(nbr)
    lib = _lib.load()
    dev = nbr.device
    order = torch.empty(n, dtype=torch.int32, device=dev)
    nbr_sorted = torch.empty_like(nbr)
    blk = torch.empty((kvol, (n + 31) // 32), dtype=torch.uint8, device=dev)
    ws_bytes = lib.pv2_rulebook_row_order_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with _lib.on_device(dev):
        _lib.check(lib.pv2_rulebook_row_order(_lib.ptr(nbr), n, kvol, _lib.ptr(order), _lib.ptr(nbr_sorted),
                                              _lib.ptr(blk), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
                   "pv2_rulebook_row_order")
    return TileMap(nbr_sorted, order, blk)


def build_row_order(nbr: torch.Tensor) -> Optional[torch.Tensor]:
    """order[pos] = row (stable sort of the rows by neighbour-presence mask), or None when no order is used."""
    return build_tile_map(nbr).order


class SubMRulebook:
    """nbr[k][j] = input row feeding output row j through kernel offset k, or -1; `order` groups rows into tiles."""

    def __init__(self, nbr: torch.Tensor, ksize: int, pair_count: Optional[torch.Tensor]):
        self.nbr = nbr                    # canonical map [K, n] (natural row order; what the parity tests compare)
        self.ksize = ksize
        self._pair_count = pair_count
        self.tmap = build_tile_map(nbr)   # what the kernels consume

    @property
    def order(self):
        return self.tmap.order

    @property
    def num_pairs(self) -> int:
        return int(self._pair_count.item())


class DownRulebook:
    """Kernel-2 stride-2 rulebook shared by SparseConv3d (fine->coarse) and SparseInverseConv3d (coarse->fine)."""

    def __init__(self, in_indices, in_shape, out_indices, out_shape, in2out, koff, nbr_down, nbr_up):
        self.in_indices = in_indices
        self.in_shape = list(in_shape)
        self.out_indices = out_indices
        self.out_shape = list(out_shape)
        self.in2out = in2out
        self.koff = koff
        self.nbr_down = nbr_down  # [8, n_out]
        self.nbr_up = nbr_up      # [8, n_in]
        self.tmap_down = build_tile_map(nbr_down)   # coarse rows as outputs
        self.tmap_up = build_tile_map(nbr_up)       # fine rows as outputs (one offset per row -> 8x fewer chunks)


def _check_indices(indices: torch.Tensor) -> None:
    if not indices.is_cuda:
        raise RuntimeError("ponderv2_b200.spconv: indices must live on a CUDA device (no CPU fallback)")
    if indices.dtype != torch.int32 or indices.dim() != 2 or indices.shape[1] != 4:
        raise ValueError("indices must be int32 [N, 4] (batch, c0, c1, c2)")


def build_subm_rulebook(indices: torch.Tensor, spatial_shape: Sequence[int], ksize: int,
                        count_pairs: bool = True) -> SubMRulebook:
    _check_indices(indices)
    indices = indices.contiguous()
    n = indices.shape[0]
    lib = _lib.load()
    nbr = torch.empty((ksize ** 3, n), dtype=torch.int32, device=indices.device)
    ws_bytes = lib.pv2_rulebook_workspace_bytes(n)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=indices.device)
    pc = torch.zeros(1, dtype=torch.int64, device=indices.device) if count_pairs else None
    with _lib.on_device(indices.device):
        _lib.check(lib.pv2_rulebook_subm(_lib.ptr(indices), n, _lib.i32x3(spatial_shape), ksize, _lib.ptr(nbr),
                                         _lib.ptr(pc), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
                   "pv2_rulebook_subm")
    return SubMRulebook(nbr, ksize, pc)


def build_down_rulebook(indices: torch.Tensor, spatial_shape: Sequence[int]) -> DownRulebook:
    _check_indices(indices)
    indices = indices.contiguous()
    n = indices.shape[0]
    dev = indices.device
    lib = _lib.load()
    out_coords = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    in2out = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    koff = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    n_out_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.pv2_rulebook_workspace_bytes(n)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with _lib.on_device(dev):
        _lib.check(lib.pv2_rulebook_down(_lib.ptr(indices), n, _lib.i32x3(spatial_shape), _lib.ptr(out_coords),
                                         _lib.ptr(in2out), _lib.ptr(koff), _lib.ptr(n_out_dev), _lib.ptr(ws),
                                         ws_bytes, _lib.stream_ptr()), "pv2_rulebook_down")
        n_out = int(n_out_dev.item())  # the one host sync of the strided rulebook (tensor shapes depend on it)
        nbr_down = torch.empty((8, n_out), dtype=torch.int32, device=dev)
        nbr_up = torch.empty((8, n), dtype=torch.int32, device=dev)
        _lib.check(lib.pv2_rulebook_down_maps(_lib.ptr(in2out), _lib.ptr(koff), n, n_out, _lib.ptr(nbr_down),
                                              _lib.ptr(nbr_up), _lib.stream_ptr()), "pv2_rulebook_down_maps")
    out_shape = [(int(s) - 2) // 2 + 1 for s in spatial_shape]
    return DownRulebook(indices, spatial_shape, out_coords[:n_out], out_shape, in2out[:n], koff[:n], nbr_down, nbr_up)


# ----------------------------------------------------------------------------------------------
# arithmetic
# ----------------------------------------------------------------------------------------------
def _gather_gemm(x: torch.Tensor, w3: torch.Tensor, bias: Optional[torch.Tensor], tmap: TileMap,
                 n_out: int) -> torch.Tensor:
    """y[j] = bias + sum_k w3[:, k, :] @ x[nbr[k][j]];  w3 is [Cout, K, Cin] (any strides with unit Cin stride)."""
    nbr, order = tmap.nbr, tmap.order
    lib = _lib.load()
    cout, kvol, cin = w3.shape
    assert w3.stride(2) == 1
    x = x.contiguous()
    y = torch.empty((n_out, cout), dtype=x.dtype, device=x.device)
    b = x.element_size()
    # algorithmic bytes: read X once, write Y once, the weights, the int32 neighbour map
    nbytes = x.shape[0] * cin * b + n_out * cout * b + kvol * cin * cout * b + 4 * kvol * n_out
    dcode = _lib.dtype_code(x.dtype)
    ws_bytes = lib.pv2_spconv_workspace_bytes(x.shape[0], cin, cout, kvol, dcode)
    with _lib.on_device(x.device), _lib.timed("pv2_spconv_gather_gemm", nbytes, 0):
        ws = _lib.workspace(ws_bytes, x.device)
        _lib.check(lib.pv2_spconv_gather_gemm(_lib.ptr(x), _lib.C.c_void_p(w3.data_ptr()), w3.stride(0), w3.stride(1),
                                              _lib.ptr(bias), _lib.ptr(nbr), _lib.ptr(order), _lib.ptr(y), x.shape[0],
                                              n_out, cin,
                                              cout, kvol, dcode, _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
                   "pv2_spconv_gather_gemm")
    return y


_WGRAD_STREAMS = {}


def _wgrad_stream(dev: torch.device) -> "torch.cuda.Stream":
    """Side stream for weight gradients written straight into a flat gradient buffer (see _SparseConvFunction.backward)."""
    key = (dev.type, dev.index)
    if key not in _WGRAD_STREAMS:
        _WGRAD_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _WGRAD_STREAMS[key]


def _wgrad(x: torch.Tensor, dy: torch.Tensor, tmap: TileMap, kvol: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dw[co, k, ci] += sum_j dy[j, co] x[nbr[k][j], ci].  `out` (fp32 [cout, kvol, cin], contiguous) is accumulated into;
    without it a zeroed tensor is allocated."""
    nbr, order = tmap.nbr, tmap.order
    lib = _lib.load()
    cin, cout = x.shape[1], dy.shape[1]
    dw = out if out is not None else torch.zeros((cout, kvol, cin), dtype=torch.float32, device=x.device)
    b = x.element_size()
    nbytes = x.shape[0] * cin * b + dy.shape[0] * cout * b + kvol * cin * cout * 4 + 4 * kvol * dy.shape[0]
    ws_bytes = lib.pv2_wgrad_workspace_bytes(x.shape[0], dy.shape[0], cin, cout) if x.dtype == torch.float32 else 0
    with _lib.on_device(x.device), _lib.timed("pv2_spconv_wgrad", nbytes, 0):
        ws = _lib.workspace(ws_bytes, x.device)
        _lib.check(lib.pv2_spconv_wgrad(_lib.ptr(x.contiguous()), _lib.ptr(dy.contiguous()), _lib.ptr(nbr),
                                        _lib.ptr(order), _lib.ptr(tmap.blk_active), _lib.ptr(dw), x.shape[0],
                                        dy.shape[0], cin, cout, kvol,
                                        _lib.dtype_code(x.dtype), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
                   "pv2_spconv_wgrad")
    return dw


class _SparseConvFunction(torch.autograd.Function):
    """Differentiable gather-GEMM.  nbr_fwd is [K, n_out] (rows index x), nbr_bwd is [K, n_in] (rows index dy);
    `flip` mirrors the kernel offsets for the data gradient (submanifold symmetry: nbr[k][j]=i <=> nbr[K-1-k][i]=j)."""

    @staticmethod
    def forward(ctx, x, weight, bias, map_fwd: TileMap, map_bwd: TileMap, n_out: int, flip: bool):
        cout, cin = weight.shape[0], weight.shape[-1]
        ctx.cin_orig = cin
        if cin % 8 != 0:
            # ragged stems (6 colour+normal / 4 xyz+strength channels): zero-pad to 8 so that the rows are whole 16-byte
            # pieces for the tensor-core kernels; the padded weight columns never see a gradient
            pad = 8 - cin % 8
            x = torch.nn.functional.pad(x, (0, pad))
            weight = torch.nn.functional.pad(weight, (0, pad))
            cin += pad
        compute_dtype = x.dtype
        if torch.is_autocast_enabled():
            compute_dtype = torch.get_autocast_dtype("cuda")
            if compute_dtype == torch.float16:  # B200 path computes in bf16 where the reference used fp16 autocast
                compute_dtype = torch.bfloat16
        xc = x.to(compute_dtype)
        w3 = weight.reshape(cout, -1, cin).to(compute_dtype)
        b = bias.float() if bias is not None else None
        y = _gather_gemm(xc, w3, b, map_fwd, n_out)
        # A parameter re-homed by ponderv2_b200.dist.FlatParameters carries a view of the flat fp32 gradient buffer: the
        # weight-gradient kernel then accumulates straight into it (no zero-filled temporary, no autograd accumulation
        # kernel) on a side stream, off the critical path of the data gradients.
        sink = getattr(weight, "_pv2_sink", None)
        ctx.sink = sink if (sink is not None and cin == ctx.cin_orig and sink[1].dtype == torch.float32
                            and sink[1].is_contiguous() and sink[1].numel() == weight.numel()) else None
        ctx.param = weight if ctx.sink is not None else None
        ctx.save_for_backward(xc, w3)
        ctx.map_fwd, ctx.map_bwd = map_fwd, map_bwd
        ctx.flip = flip
        ctx.has_bias = bias is not None
        ctx.weight_shape = weight.shape   # padded shape when cin was ragged; sliced back in backward
        ctx.weight_dtype = weight.dtype
        ctx.x_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, w3 = ctx.saved_tensors
        dy = dy.contiguous().to(xc.dtype)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            cout_, kvol_, cin_ = w3.shape
            wt = torch.empty((cin_, kvol_, cout_), dtype=w3.dtype, device=w3.device)   # [Cin, K, Cout], k-flipped for SubM
            lib = _lib.load()
            with _lib.on_device(w3.device):
                _lib.check(lib.pv2_spconv_dgrad_weights(_lib.ptr(w3.contiguous()), _lib.ptr(wt), cout_, kvol_, cin_,
                                                        int(ctx.flip), _lib.dtype_code(w3.dtype), _lib.stream_ptr()),
                           "pv2_spconv_dgrad_weights")
            dx = _gather_gemm(dy, wt, None, ctx.map_bwd, xc.shape[0]).to(ctx.x_dtype)[:, :ctx.cin_orig]
        if ctx.needs_input_grad[1] and ctx.sink is not None:
            flat, view = ctx.sink
            cur = torch.cuda.current_stream(dy.device)
            side = _wgrad_stream(dy.device)
            side.wait_stream(cur)                       # dy (and, earlier, xc) are produced on the launching stream
            with torch.cuda.stream(side):
                _wgrad(xc, dy, ctx.map_fwd, w3.shape[1], out=view.view(w3.shape[0], w3.shape[1], w3.shape[2]))
            xc.record_stream(side); dy.record_stream(side)
            flat.note_aux_stream(side)                  # optimizer / all-reduce wait for it
            flat.mark_ready(ctx.param)
        elif ctx.needs_input_grad[1]:
            dw = _wgrad(xc, dy, ctx.map_fwd, w3.shape[1]).reshape(ctx.weight_shape).to(ctx.weight_dtype)[..., :ctx.cin_orig]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(0)
        return dx, dw, db, None, None, None, None


# ----------------------------------------------------------------------------------------------
# module surface
# ----------------------------------------------------------------------------------------------
class SparseConvTensor:
    def __init__(self, features: torch.Tensor, indices: torch.Tensor, spatial_shape: Union[List[int], Sequence[int]],
                 batch_size: int, grid=None, voxel_num=None, indice_dict: Optional[dict] = None, benchmark=False):
        _check_indices(indices)
        if features.dim() != 2 or features.shape[0] != indices.shape[0]:
            raise ValueError("features must be [N, C] with one row per index row")
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark

    def replace_feature(self, new_features: torch.Tensor) -> "SparseConvTensor":
        out = SparseConvTensor(new_features, self.indices, self.spatial_shape, self.batch_size, self.grid,
                               self.voxel_num, self.indice_dict, self.benchmark)
        return out

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key, None)

    @property
    def spatial_size(self) -> int:
        return int(math.prod(self.spatial_shape))

    def dense(self, channels_first: bool = True) -> torch.Tensor:
        shape = [self.batch_size, *self.spatial_shape, self.features.shape[1]]
        out = torch.zeros(shape, dtype=self.features.dtype, device=self.features.device)
        idx = self.indices.long()
        out[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = self.features
        return out.permute(0, 4, 1, 2, 3).contiguous() if channels_first else out


class SparseModule(nn.Module):
    """Marker base class: modules that consume/produce SparseConvTensor."""


def _is_sparse_module(m: nn.Module) -> bool:
    return isinstance(m, SparseModule)


class SparseSequential(SparseModule):
    """Sequential container; plain nn.Modules are applied to `.features`."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError(f"index {idx} is out of range")
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        for module in self._modules.values():
            if _is_sparse_module(module):
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input = input.replace_feature(module(input.features))
            else:
                input = module(input)
        return input


class Identity(SparseModule):
    def forward(self, input):
        return input


def _triple(v) -> List[int]:
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return [int(a) for a in v]
    return [int(v)] * 3


class _SparseConvBase(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, **kwargs):
        super().__init__()
        if groups != 1 or _triple(dilation) != [1, 1, 1]:
            raise NotImplementedError("ponderv2_b200.spconv: groups/dilation are not used by the reference path")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _triple(kernel_size)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        # spconv's default: kaiming-uniform(a=sqrt(5)) over fan_in = K*Cin, bias U(+-1/sqrt(fan_in))
        fan_in = self.in_channels * math.prod(self.kernel_size)
        gain = math.sqrt(2.0 / (1 + 5.0))
        bound = gain * math.sqrt(3.0 / fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                b = 1.0 / math.sqrt(fan_in)
                self.bias.uniform_(-b, b)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"indice_key={self.indice_key}")


class SubMConv3d(_SparseConvBase):
    def forward(self, input: SparseConvTensor) -> SparseConvTensor:
        ks = self.kernel_size
        if not (ks[0] == ks[1] == ks[2] and ks[0] in (1, 3, 5)):
            raise NotImplementedError("SubMConv3d: cubic kernels of size 1, 3 or 5 only")
        n = input.features.shape[0]
        if ks[0] == 1:
            key = ("__identity__", n)
            rb = input.indice_dict.get(key)
            if rb is None:
                nbr = torch.arange(n, dtype=torch.int32, device=input.features.device).view(1, n)
                rb = SubMRulebook(nbr, 1, None)
                input.indice_dict[key] = rb
        else:
            rb = input.find_indice_pair(self.indice_key)
            if rb is None:
                rb = build_subm_rulebook(input.indices, input.spatial_shape, ks[0])
                if self.indice_key is not None:
                    input.indice_dict[self.indice_key] = rb
            elif not isinstance(rb, SubMRulebook) or rb.ksize != ks[0] or rb.nbr.shape[1] != n:
                raise ValueError(f"indice_key {self.indice_key!r} holds a rulebook of a different conv")
        y = _SparseConvFunction.apply(input.features, self.weight, self.bias, rb.tmap, rb.tmap, n, True)
        return input.replace_feature(y)


class SparseConv3d(_SparseConvBase):
    def forward(self, input: SparseConvTensor) -> SparseConvTensor:
        if self.kernel_size != [2, 2, 2] or self.stride != [2, 2, 2] or self.padding != [0, 0, 0]:
            raise NotImplementedError("SparseConv3d: only kernel_size=2, stride=2, padding=0 (the reference's use)")
        rb = input.find_indice_pair(self.indice_key)
        if rb is None:
            rb = build_down_rulebook(input.indices, input.spatial_shape)
            if self.indice_key is not None:
                input.indice_dict[self.indice_key] = rb
        n_out = rb.out_indices.shape[0]
        y = _SparseConvFunction.apply(input.features, self.weight, self.bias, rb.tmap_down, rb.tmap_up, n_out, False)
        return SparseConvTensor(y, rb.out_indices, rb.out_shape, input.batch_size, input.grid, input.voxel_num,
                                input.indice_dict, input.benchmark)


class SparseInverseConv3d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, indice_key=indice_key, **kwargs)

    def forward(self, input: SparseConvTensor) -> SparseConvTensor:
        rb = input.find_indice_pair(self.indice_key)
        if not isinstance(rb, DownRulebook):
            raise ValueError(f"SparseInverseConv3d: no SparseConv3d rulebook stored under {self.indice_key!r}")
        if self.kernel_size != [2, 2, 2]:
            raise NotImplementedError("SparseInverseConv3d: kernel_size=2 only")
        n_fine = rb.in_indices.shape[0]
        y = _SparseConvFunction.apply(input.features, self.weight, self.bias, rb.tmap_up, rb.tmap_down, n_fine, False)
        return SparseConvTensor(y, rb.in_indices, rb.in_shape, input.batch_size, input.grid, input.voxel_num,
                                input.indice_dict, input.benchmark)
