"""SpUNet-v1m3 (prompt-driven normalisation) on the B200 sparse-conv kernels.

Host-side mirror of ponder/models/sparse_unet/spconv_unet_v1m3_pdnorm.py:23-460 (`PDBatchNorm`, `BasicBlock`,
`SPConvDown/Up/PatchEmbedding`, `SpUNetBase` registered as "SpUNet-v1m3"): same constructor arguments, forward contract
(`input_dict{grid_coord, feat, offset, condition, [context]} -> [N, channels[-1]]`) and parameter / buffer names
(`bns.{i}.running_mean`, `modulation.1.weight`, `proj_conv`, `proj_norm`, ... pinned by
tests/golden/spunet_v1m3_state.json), so PPT checkpoints load unchanged.

What differs is where the arithmetic runs.  The reference normalises ([N, C] pass), then modulates with the context
(`feat * (1 + scale) + shift`, two more passes), then ReLU / residual add (one or two more).  Scale and shift are
per-CHANNEL (the context is one row per batch), so the modulation folds into the BatchNorm's affine pair and the whole
chain is the fused bn_act kernel pair (bn_act.bn_act_modulated); the only extra device work per norm layer is the
[1, 256] x [256, 2C] modulation Linear.
"""
from __future__ import annotations

from collections import OrderedDict
from functools import partial

import torch
from torch import nn

from .backbone import make_sparse_indices, prebuild_rulebooks
from .bn_act import bn_act_modulated
from .spconv import pytorch as spconv


class PDBatchNorm(nn.Module):
    """spconv_unet_v1m3_pdnorm.py:23-72.  `forward(feat, condition, context)` keeps the reference's signature;
    `fused(...)` adds the residual / ReLU that follow it in every block."""

    def __init__(self, num_features, context_channels=256, eps=1e-3, momentum=0.01,
                 conditions=("ScanNet", "S3DIS", "Structured3D"), decouple=True, adaptive=False, affine=True):
        super().__init__()
        self.conditions = tuple(conditions)
        self.decouple, self.adaptive, self.affine = decouple, adaptive, affine
        if decouple:
            self.bns = nn.ModuleList([nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine)
                                      for _ in self.conditions])
        else:
            self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine)
        if adaptive:
            self.modulation = nn.Sequential(nn.SiLU(), nn.Linear(context_channels, 2 * num_features, bias=True))

    def fused(self, feat, condition=None, context=None, residual=None, relu=False):
        if self.decouple:
            if condition not in self.conditions:
                raise ValueError(f"PDBatchNorm: condition {condition!r} not in {self.conditions}")
            bn = self.bns[self.conditions.index(condition)]
        else:
            bn = self.bn
        scale = shift = None
        if self.adaptive:
            if context is None:
                raise ValueError("PDBatchNorm(adaptive=True) needs a context row")
            if context.shape[0] != 1:
                raise NotImplementedError("PDBatchNorm: one context row per batch (as PonderIndoor.extract_feature builds it)")
            shift, scale = self.modulation(context.float()).chunk(2, dim=1)
        return bn_act_modulated(feat, bn, scale, shift, residual, relu)

    def forward(self, feat, condition=None, context=None):
        return self.fused(feat, condition, context)


class BasicBlock(spconv.SparseModule):
    """:75-143; input and output are the reference's (x, condition, context) triples."""

    expansion = 1

    def __init__(self, in_channels, embed_channels, stride=1, norm_fn=None, indice_key=None, bias=False):
        super().__init__()
        assert norm_fn is not None
        self.in_channels, self.embed_channels = in_channels, embed_channels
        if in_channels == embed_channels:
            self.proj = spconv.SparseSequential(nn.Identity())
        else:
            self.proj_conv = spconv.SubMConv3d(in_channels, embed_channels, kernel_size=1, bias=False)
            self.proj_norm = norm_fn(embed_channels)
        self.conv1 = spconv.SubMConv3d(in_channels, embed_channels, kernel_size=3, stride=stride, padding=1, bias=bias,
                                       indice_key=indice_key)
        self.bn1 = norm_fn(embed_channels)
        self.relu = nn.ReLU()
        self.conv2 = spconv.SubMConv3d(embed_channels, embed_channels, kernel_size=3, stride=stride, padding=1,
                                       bias=bias, indice_key=indice_key)
        self.bn2 = norm_fn(embed_channels)
        self.stride = stride

    def forward(self, x):
        x, condition, context = x
        out = self.conv1(x)
        out = out.replace_feature(self.bn1.fused(out.features, condition, context, None, True))
        out = self.conv2(out)
        if self.in_channels == self.embed_channels:
            res = x.features
        else:
            res = self.proj_norm.fused(self.proj_conv(x).features, condition, context, None, False)
        out = out.replace_feature(self.bn2.fused(out.features, condition, context, res, True))
        return out, condition, context


class _ConvNormReLU(nn.Module):
    def forward(self, x):
        x, condition, context = x
        out = self.conv(x)
        return out.replace_feature(self.bn.fused(out.features, condition, context, None, True))


class SPConvDown(_ConvNormReLU):
    def __init__(self, in_channels, out_channels, indice_key, kernel_size=2, bias=False, norm_fn=None):
        super().__init__()
        self.conv = spconv.SparseConv3d(in_channels, out_channels, kernel_size=kernel_size, stride=kernel_size, bias=bias,
                                        indice_key=indice_key)
        self.bn = norm_fn(out_channels)
        self.relu = nn.ReLU()


class SPConvUp(_ConvNormReLU):
    def __init__(self, in_channels, out_channels, indice_key, kernel_size=2, bias=False, norm_fn=None):
        super().__init__()
        self.conv = spconv.SparseInverseConv3d(in_channels, out_channels, kernel_size=kernel_size, bias=bias,
                                               indice_key=indice_key)
        self.bn = norm_fn(out_channels)
        self.relu = nn.ReLU()


class SPConvPatchEmbedding(_ConvNormReLU):
    def __init__(self, in_channels, out_channels, kernel_size=5, norm_fn=None):
        super().__init__()
        self.conv = spconv.SubMConv3d(in_channels, out_channels, kernel_size=kernel_size, padding=1, bias=False,
                                      indice_key="stem")
        self.bn = norm_fn(out_channels)
        self.relu = nn.ReLU()


class _TripleSequential(nn.Sequential):
    """The reference runs its residual stacks through spconv.SparseSequential, whose non-sparse branch passes the
    (x, condition, context) list from block to block; this does the same and keeps the `block{i}` child names."""

    def forward(self, x):
        for m in self:
            x = m(x)
        return x


class SpUNetPDNorm(nn.Module):
    """:246-460, registered as "SpUNet-v1m3"."""

    def __init__(self, in_channels, num_classes=0, base_channels=32, context_channels=256,
                 channels=(32, 64, 128, 256, 256, 128, 96, 96), layers=(2, 3, 4, 6, 2, 2, 2, 2), cls_mode=False,
                 conditions=("ScanNet", "S3DIS", "Structured3D"), zero_init=True, norm_decouple=True, norm_adaptive=True,
                 norm_affine=False):
        super().__init__()
        assert len(layers) % 2 == 0 and len(layers) == len(channels)
        if cls_mode:
            raise NotImplementedError("cls_mode is a fine-tuning option outside the pretraining path")
        self.in_channels, self.num_classes, self.base_channels = in_channels, num_classes, base_channels
        self.channels, self.layers = tuple(channels), tuple(layers)
        self.num_stages = len(layers) // 2
        self.cls_mode, self.conditions, self.zero_init = cls_mode, tuple(conditions), zero_init
        norm_fn = partial(PDBatchNorm, eps=1e-3, momentum=0.01, conditions=conditions, context_channels=context_channels,
                          decouple=norm_decouple, adaptive=norm_adaptive, affine=norm_affine)
        self.conv_input = SPConvPatchEmbedding(in_channels, base_channels, kernel_size=5, norm_fn=norm_fn)
        self.down, self.up, self.enc, self.dec = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        enc_c, dec_c = base_channels, channels[-1]
        nch = len(channels)
        for s in range(self.num_stages):
            self.down.append(SPConvDown(enc_c, channels[s], kernel_size=2, bias=False, indice_key=f"spconv{s + 1}",
                                        norm_fn=norm_fn))
            self.enc.append(_TripleSequential(OrderedDict(
                (f"block{i}", BasicBlock(channels[s], channels[s], norm_fn=norm_fn, indice_key=f"subm{s + 1}"))
                for i in range(layers[s]))))
            self.up.append(SPConvUp(channels[nch - s - 2], dec_c, kernel_size=2, bias=False,
                                    indice_key=f"spconv{s + 1}", norm_fn=norm_fn))
            self.dec.append(_TripleSequential(OrderedDict(
                (f"block{i}", BasicBlock(dec_c + enc_c if i == 0 else dec_c, dec_c, norm_fn=norm_fn,
                                         indice_key=f"subm{s}"))
                for i in range(layers[nch - s - 1]))))
            enc_c, dec_c = channels[s], channels[nch - s - 2]
        self.final = (spconv.SubMConv3d(channels[-1], num_classes, kernel_size=1, padding=1, bias=True)
                      if num_classes > 0 else spconv.Identity())
        self.apply(self._init_weights)

    def _init_weights(self, m):
        # :400-416; children are visited before their parents, so the modulation Linear is first trunc_normal'd and
        # then zeroed by its PDBatchNorm when zero_init
        if isinstance(m, (nn.Linear, spconv.SubMConv3d)):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm1d):
            if m.affine:
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, PDBatchNorm):
            if self.zero_init and m.adaptive:
                nn.init.constant_(m.modulation[-1].weight, 0)
                nn.init.constant_(m.modulation[-1].bias, 0)

    def forward(self, input_dict):
        grid_coord, feat, offset = input_dict["grid_coord"], input_dict["feat"], input_dict["offset"]
        if not grid_coord.is_cuda:
            raise RuntimeError("SpUNet-v1m3: inputs must be CUDA tensors (ponderv2_b200 has no CPU path)")
        condition = input_dict["condition"][0]
        context = input_dict.get("context")
        shape = input_dict.get("sparse_shape")
        if shape is None:
            shape = torch.add(torch.max(grid_coord, dim=0).values, 96).tolist()
        x = spconv.SparseConvTensor(features=feat, indices=make_sparse_indices(grid_coord, offset),
                                    spatial_shape=shape, batch_size=int(offset.shape[0]))
        prebuild_rulebooks(x, self.conv_input.conv.kernel_size[0], self.num_stages)
        x = self.conv_input([x, condition, context])
        skips = [x]
        for s in range(self.num_stages):
            x = self.down[s]([x, condition, context])
            x, _, _ = self.enc[s]([x, condition, context])
            skips.append(x)
        x = skips.pop(-1)
        for s in reversed(range(self.num_stages)):
            x = self.up[s]([x, condition, context])
            skip = skips.pop(-1)
            x = x.replace_feature(torch.cat((x.features, skip.features), dim=1))
            x, _, _ = self.dec[s]([x, condition, context])
        return self.final(x).features
