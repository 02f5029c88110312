"""SpUNet-v1m1 voxel backbone on the B200 sparse-conv kernels.

Host-side mirror of ponder/models/sparse_unet/spconv_unet_v1m1_base.py:21-278 (`BasicBlock`, `SpUNetBase`): same
constructor arguments, same forward contract (`input_dict{grid_coord, feat, offset} -> [N, channels[-1]]`) and the
same parameter / buffer names, so reference checkpoints load unchanged (hooks/misc.py:208-253; the name+shape
contract is pinned by tests/golden/spunet_v1m1_state.json).  Everything below `.features` is torch plumbing
(BatchNorm1d, ReLU, cat); the convolutions and rulebooks are ponderv2_b200.spconv -> libpv2_b200.
"""
from __future__ import annotations

from collections import OrderedDict
from functools import partial

import torch
from torch import nn

from . import _lib
from .bn_act import bn_act
from .spconv import pytorch as spconv


def make_sparse_indices(grid_coord: torch.Tensor, offset: torch.Tensor) -> torch.Tensor:
    """[N,4] int32 (batch, c0, c1, c2) from `grid_coord` [N,3] and cumulative `offset` [B]
    (offset2batch + cat, spconv_unet_v1m1_base.py:247-253), in one kernel and without the reference's
    CPU list-comprehension round trip (ponder/models/utils.py:11-26)."""
    lib = _lib.load()
    n = grid_coord.shape[0]
    gc = grid_coord.contiguous().long()
    off = offset.contiguous().long()
    out = torch.empty((n, 4), dtype=torch.int32, device=gc.device)
    with _lib.on_device(gc.device):
        _lib.check(lib.pv2_make_indices(_lib.ptr(gc), _lib.ptr(off), n, off.shape[0], _lib.ptr(out),
                                        _lib.stream_ptr()), "pv2_make_indices")
    return out


class ResidualBlock(spconv.SparseModule):
    """conv3-bn-relu-conv3-bn (+ 1x1 projected residual) - relu; mirrors BasicBlock (:21-83)."""

    expansion = 1

    def __init__(self, in_channels, embed_channels, stride=1, norm_fn=None, indice_key=None, bias=False):
        super().__init__()
        assert norm_fn is not None
        if in_channels == embed_channels:
            self.proj = spconv.SparseSequential(nn.Identity())
        else:
            self.proj = spconv.SparseSequential(
                spconv.SubMConv3d(in_channels, embed_channels, kernel_size=1, bias=False), norm_fn(embed_channels))
        self.conv1 = spconv.SubMConv3d(in_channels, embed_channels, kernel_size=3, stride=stride, padding=1,
                                       bias=bias, indice_key=indice_key)
        self.bn1 = norm_fn(embed_channels)
        self.relu = nn.ReLU()
        self.conv2 = spconv.SubMConv3d(embed_channels, embed_channels, kernel_size=3, stride=stride, padding=1,
                                       bias=bias, indice_key=indice_key)
        self.bn2 = norm_fn(embed_channels)
        self.stride = stride

    def forward(self, x):
        # conv - bn - relu - conv - bn - (+ residual) - relu, with each bn(+add)+relu a fused kernel pair (bn_act.py)
        out = self.conv1(x)
        out = out.replace_feature(bn_act(out.features, self.bn1, None, True))
        out = self.conv2(out)
        if len(self.proj) == 2:   # 1x1 projection + its own BatchNorm (no ReLU)
            res = bn_act(self.proj[0](x).features, self.proj[1], None, False)
        else:
            res = x.features
        return out.replace_feature(bn_act(out.features, self.bn2, res, True))


class ConvBNReLU(spconv.SparseSequential):
    """SparseSequential(conv, BatchNorm1d, ReLU) — same children / state_dict keys ('0', '1', '2') as the reference's
    blocks (spconv_unet_v1m1_base.py:111-119,134-145,170-180) — with the BatchNorm + ReLU pair run as one fused op."""

    def forward(self, input):
        out = self[0](input)
        return out.replace_feature(bn_act(out.features, self[1], None, True))


def prebuild_rulebooks(x: "spconv.SparseConvTensor", stem_kernel: int, num_stages: int) -> None:
    """All coordinate work of the step up front.  The rulebooks depend on coordinates only, and each strided one
    ends in a host read of its output count (tensor shapes depend on it); built lazily inside the layer stack
    those four syncs drain a full launch queue each.  Built here, nothing is queued behind them yet and the host
    runs ahead of the GPU for the rest of the forward pass.  Keys are the reference's indice_keys
    (spconv_unet_v1m1_base.py:111-177): stem (k5), subm0..4 (k3), spconv1..4 (k2 s2)."""
    d = x.indice_dict
    ind, shape = x.indices, x.spatial_shape
    if "stem" not in d:
        d["stem"] = spconv.build_subm_rulebook(ind, shape, stem_kernel, count_pairs=False)
    if "subm0" not in d:
        d["subm0"] = spconv.build_subm_rulebook(ind, shape, 3, count_pairs=False)
    for s in range(num_stages):
        key = f"spconv{s + 1}"
        if key not in d:
            d[key] = spconv.build_down_rulebook(ind, shape)
        ind, shape = d[key].out_indices, d[key].out_shape
        if f"subm{s + 1}" not in d:
            d[f"subm{s + 1}"] = spconv.build_subm_rulebook(ind, shape, 3, count_pairs=False)


class SpUNetBase(nn.Module):
    def __init__(self, in_channels, num_classes, base_channels=32, channels=(32, 64, 128, 256, 256, 128, 96, 96),
                 layers=(2, 3, 4, 6, 2, 2, 2, 2), cls_mode=False):
        super().__init__()
        assert len(layers) % 2 == 0 and len(layers) == len(channels)
        if cls_mode:
            raise NotImplementedError("cls_mode is a fine-tuning option outside the pretraining path")
        self.in_channels, self.num_classes, self.base_channels = in_channels, num_classes, base_channels
        self.channels, self.layers = tuple(channels), tuple(layers)
        self.num_stages = len(layers) // 2
        self.cls_mode = cls_mode
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)

        self.conv_input = ConvBNReLU(
            spconv.SubMConv3d(in_channels, base_channels, kernel_size=5, padding=1, bias=False, indice_key="stem"),
            norm_fn(base_channels), nn.ReLU())
        self.down, self.up, self.enc, self.dec = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        enc_c, dec_c = base_channels, channels[-1]
        nch = len(channels)
        for s in range(self.num_stages):
            self.down.append(ConvBNReLU(
                spconv.SparseConv3d(enc_c, channels[s], kernel_size=2, stride=2, bias=False,
                                    indice_key=f"spconv{s + 1}"),
                norm_fn(channels[s]), nn.ReLU()))
            self.enc.append(spconv.SparseSequential(OrderedDict(
                (f"block{i}", ResidualBlock(channels[s], channels[s], norm_fn=norm_fn, indice_key=f"subm{s + 1}"))
                for i in range(layers[s]))))
            self.up.append(ConvBNReLU(
                spconv.SparseInverseConv3d(channels[nch - s - 2], dec_c, kernel_size=2, bias=False,
                                           indice_key=f"spconv{s + 1}"),
                norm_fn(dec_c), nn.ReLU()))
            self.dec.append(spconv.SparseSequential(OrderedDict(
                (f"block{i}", ResidualBlock(dec_c + enc_c if i == 0 else dec_c, dec_c, norm_fn=norm_fn,
                                            indice_key=f"subm{s}"))
                for i in range(layers[nch - s - 1]))))
            enc_c, dec_c = channels[s], channels[nch - s - 2]
        self.final = (spconv.SubMConv3d(channels[-1], num_classes, kernel_size=1, padding=1, bias=True)
                      if num_classes > 0 else spconv.Identity())
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        # spconv_unet_v1m1_base.py:228-240 (timm trunc_normal_ == torch.nn.init.trunc_normal_ with a=-2, b=2)
        if isinstance(m, (nn.Linear, spconv.SubMConv3d)):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _prebuild_rulebooks(self, x: "spconv.SparseConvTensor") -> None:
        prebuild_rulebooks(x, self.conv_input[0].kernel_size[0], self.num_stages)

    def forward(self, input_dict):
        grid_coord, feat, offset = input_dict["grid_coord"], input_dict["feat"], input_dict["offset"]
        if not grid_coord.is_cuda:
            raise RuntimeError("SpUNetBase: inputs must be CUDA tensors (ponderv2_b200 has no CPU path)")
        shape = input_dict.get("sparse_shape")
        if shape is None:  # one host sync, as in the reference (:248)
            shape = torch.add(torch.max(grid_coord, dim=0).values, 96).tolist()
        x = spconv.SparseConvTensor(features=feat, indices=make_sparse_indices(grid_coord, offset),
                                    spatial_shape=shape, batch_size=int(offset.shape[0]))
        self._prebuild_rulebooks(x)
        x = self.conv_input(x)
        skips = [x]
        for s in range(self.num_stages):
            x = self.enc[s](self.down[s](x))
            skips.append(x)
        x = skips.pop(-1)
        for s in reversed(range(self.num_stages)):
            x = self.up[s](x)
            skip = skips.pop(-1)
            x = x.replace_feature(torch.cat((x.features, skip.features), dim=1))
            x = self.dec[s](x)
        return self.final(x).features
