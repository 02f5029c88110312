"""Registry-facing model classes (boundary B3): the reference's `MODELS` / `RENDERERS` names on the B200 path.

    "SpUNet-v1m1"        ponder/models/sparse_unet/spconv_unet_v1m1_base.py:86      -> backbone.SpUNetBase
    "SpUNet-v1m3"        ponder/models/sparse_unet/spconv_unet_v1m3_pdnorm.py:246  -> backbone_pdnorm.SpUNetPDNorm (§8f-3)
    "SimpleConv3D-v1m1"  ponder/models/ponder/unet3d.py:16                          -> pretrain.SimpleConv3D
    "UNet3D-v1m2"        ponder/models/ponder/unet3d.py:710 (Abstract3DUNet :530)   -> UNet3Dv1m2 (dense, cuDNN; §8f-1)
    "PonderIndoor-v2"    ponder/models/ponder/ponder_indoor_base.py:19              -> PonderIndoor
    "PonderOutdoor-v2"   ponder/models/ponder/ponder_outdoor_base.py:18             -> PonderOutdoor
    "NeuSModel"          ponder/models/ponder/render_utils/models/neus.py:7         -> render.NeuSModel

Same constructor arguments (config dicts with `type` keys built through the registry), same `forward(data_dict) ->
dict(loss=..., <name>_loss=...)` contract on the collate dict the reference's dataloader produces, same parameter names.
`install_into_reference()` registers these classes in the reference's own registries (force=True) and aliases
`spconv.pytorch` / `smooth_sampler`, after which the reference's configs and `ponder/engines` run unchanged on this
library.  The semantic branch (`render_semantic=True`, §8f-4) takes the class text embeddings as a
tensor / file (`class_embedding=`) or computes them with CLIP when that package is importable.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from . import rayprep
from .backbone import SpUNetBase
from .backbone_pdnorm import SpUNetPDNorm
from .pretrain import PonderIndoorStep, PonderOutdoorStep, SimpleConv3D, _SceneViews
from .render import RayBundle
from .render.neus import NeuSModel


class Registry:
    """The slice of ponder/utils/registry.py:58-315 that configs and builders use: `register_module` (decorator or
    direct, `force`), `get`, `build(cfg)` with the class name under cfg["type"]."""

    def __init__(self, name: str):
        self.name, self._modules = name, {}

    def get(self, key: str):
        return self._modules.get(key)

    def register_module(self, name=None, force: bool = False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if not force and key in self._modules:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def build(self, cfg: dict, **default_args):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise KeyError(f'cfg must be a dict with the key "type", got {cfg}')
        args = dict(cfg)
        for k, v in default_args.items():
            args.setdefault(k, v)
        t = args.pop("type")
        cls = self.get(t) if isinstance(t, str) else t
        if cls is None:
            raise KeyError(f"{t} is not in the {self.name} registry")
        return cls(**args)


MODELS = Registry("models")
RENDERERS = Registry("renderers")
MODELS.register_module("SpUNet-v1m1", module=SpUNetBase)
MODELS.register_module("SpUNet-v1m3", module=SpUNetPDNorm)
MODELS.register_module("SimpleConv3D-v1m1", module=SimpleConv3D)
RENDERERS.register_module("NeuSModel", module=NeuSModel)


def build_model(cfg: dict):
    return MODELS.build(cfg)


# ------------------------------------------------------------------------------------------------ UNet3D-v1m2 (§8f-1)
def _bcr(cin: int, cout: int) -> nn.Sequential:
    """layer order "bcr": BatchNorm3d on the INPUT channels, bias-free 3x3x3 conv, ReLU (unet3d.py:45-122); child names
    are the reference's ("batchnorm", "conv", "ReLU")."""
    m = nn.Sequential()
    m.add_module("batchnorm", nn.BatchNorm3d(cin))
    m.add_module("conv", nn.Conv3d(cin, cout, 3, padding=1, bias=False))
    m.add_module("ReLU", nn.ReLU(inplace=True))
    return m


class _Enc(nn.Module):
    def __init__(self, cin, cout, pool: bool):
        super().__init__()
        self.pooling = nn.MaxPool3d(kernel_size=(2, 2, 2)) if pool else None
        self.basic_module = _bcr(cin, cout)

    def forward(self, x):
        return self.basic_module(x if self.pooling is None else self.pooling(x))


class _Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.upsample = nn.ConvTranspose3d(cin, cout, kernel_size=3, stride=(2, 2, 2), padding=1)

    def forward(self, encoder_features, x):
        return self.upsample(x, encoder_features.size()[2:])


class _Dec(nn.Module):
    """transposed-conv upsampling to the skip's size, SUM joining, one bcr block (the SingleConv variant of
    unet3d.py:359-445)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.upsampling = _Up(cin, cout)
        self.basic_module = _bcr(cout, cout)

    def forward(self, encoder_features, x):
        return self.basic_module(encoder_features + self.upsampling(encoder_features, x))


class UNet3Dv1m2(nn.Module):
    """Dense projection network of the ScanNet / S3DIS / Structured3D configs (configs/scannet/...-base.py:26-30):
    4-level 3-D U-Net, f_maps 32..256, one bcr block per level, 1x1x1 head.  Plain cuDNN convolutions in
    channels_last_3d (not part of the hand-written hot path: SURVEY §8f-1); parameter names match the reference."""

    def __init__(self, in_channels, out_channels, final_sigmoid=False, f_maps=32, layer_order="bcr", num_groups=1,
                 num_levels=4, is_segmentation=False, **kwargs):
        super().__init__()
        if layer_order != "bcr" or is_segmentation:
            raise NotImplementedError("UNet3D-v1m2: the shipped configs use layer_order='bcr', is_segmentation=False")
        if isinstance(f_maps, int):
            f_maps = [f_maps * 2 ** k for k in range(num_levels)]
        self.encoders = nn.ModuleList([_Enc(in_channels if i == 0 else f_maps[i - 1], f, pool=i > 0)
                                       for i, f in enumerate(f_maps)])
        r = list(reversed(f_maps))
        self.decoders = nn.ModuleList([_Dec(r[i], r[i + 1]) for i in range(len(r) - 1)])
        self.final_conv = nn.Conv3d(f_maps[0], out_channels, 1)

    def forward(self, x):
        feats = []
        for enc in self.encoders:
            x = enc(x)
            feats.insert(0, x)
        for dec, skip in zip(self.decoders, feats[1:]):
            x = dec(skip, x)
        return self.final_conv(x)


MODELS.register_module("UNet3D-v1m2", module=UNet3Dv1m2)


# ------------------------------------------------------------------------------------------------ PonderIndoor-v2
def _build_projection(cfg: Optional[dict], default: dict) -> nn.Module:
    return MODELS.build(dict(cfg) if cfg is not None else default).to(memory_format=torch.channels_last_3d)


def _class_embedding(given, template, clip_model, class_name) -> torch.Tensor:
    """[n_classes, E] unit-norm text embeddings of the class prompts.  `PonderIndoor.load_semantic` (:85-118) computes
    them with CLIP at construction; here they are either handed in (`class_embedding=` tensor / array / .npy / .pt path,
    e.g. saved once from the reference) or computed the same way when the `clip` package and weights are available."""
    if given is not None:
        if isinstance(given, (str, bytes)):
            import numpy as np
            given = torch.load(given) if str(given).endswith((".pt", ".pth")) else torch.from_numpy(np.load(given))
        emb = torch.as_tensor(given).float()
        return emb / emb.norm(dim=-1, keepdim=True).clamp(min=1e-12)
    try:
        import clip                                                     # noqa: F401
    except ImportError as e:
        raise RuntimeError("PonderIndoor-v2(render_semantic=True): pass class_embedding=<[n_classes, E] text embeddings> "
                           "(CLIP is not importable here to compute them from template / class_name)") from e
    model, _ = clip.load(clip_model, device="cpu", download_root="./.cache/clip")
    multi = not isinstance(template, str)
    prompts = [t.replace("[x]", n) for n in class_name for t in (template if multi else [template])]
    with torch.no_grad():
        emb = model.encode_text(clip.tokenize(prompts)).float()
    emb = emb / emb.norm(dim=-1, keepdim=True)
    if multi:
        emb = emb.reshape(len(class_name), len(template), -1).mean(1)
        emb = emb / emb.norm(dim=-1, keepdim=True)
    return emb


class _Criteria(nn.Module):
    """`build_criteria` (ponder/models/losses/builder.py) for the PPT point loss: a list of weighted criteria; the
    shipped PPT configs use CrossEntropyLoss(loss_weight, ignore_index)."""

    def __init__(self, cfg):
        super().__init__()
        self.items = []
        for c in (cfg if isinstance(cfg, (list, tuple)) else [cfg]):
            c = dict(c)
            t = c.pop("type")
            if t != "CrossEntropyLoss":
                raise NotImplementedError(f"ppt_criteria {t}: only CrossEntropyLoss is wired")
            self.items.append((float(c.get("loss_weight", 1.0)), int(c.get("ignore_index", -1)),
                               float(c.get("label_smoothing", 0.0))))

    def forward(self, logits, target):
        return sum(w * F.cross_entropy(logits, target.long(), ignore_index=ig, label_smoothing=ls)
                   for w, ig, ls in self.items)


class PonderIndoor(PonderIndoorStep):
    """`PonderIndoor` with the reference's constructor and data_dict contract (ponder_indoor_base.py:19-706): takes the
    collate dict (coord, grid_coord, feat, offset, rgb (B,V,H,W,3), depth (B,V,H,W), intrinsic, extrinsic (B,V,4,4),
    depth_scale[, condition]) and does ray preparation on the device (ponderv2_b200.rayprep)."""

    def __init__(self, backbone, projection, renderer, mask=None, grid_shape=64, grid_size=0.02, val_ray_split=10240,
                 ray_nsample=128, padding=0.1, backbone_out_channels=96, context_channels=256, pool_type="mean",
                 render_semantic=False, conditions=None, template=None, clip_model=None, class_name=None,
                 valid_index=None, ppt_loss_weight=1.0, ppt_criteria=None, class_embedding=None, logit_scale=4.6052):
        gs = tuple(grid_shape) if isinstance(grid_shape, Sequence) else (grid_shape,) * 3
        nn.Module.__init__(self)
        if pool_type != "mean":
            raise NotImplementedError("pool_type: every shipped config uses 'mean'")
        self.val_ray_split = int(val_ray_split)
        self.backbone = MODELS.build(dict(backbone))
        self.proj_net = _build_projection(projection, dict(type="SimpleConv3D-v1m1", in_channels=96, out_channels=128))
        self.renderer = RENDERERS.build(dict(renderer))
        self.grid_shape = tuple(int(g) for g in gs)
        self.grid_size = float(grid_size)
        self.ray_nsample = int(ray_nsample)
        self.bounds = [[-0.5 - padding / 2] * 3, [0.5 + padding / 2] * 3]
        self.mask = dict(mask) if mask is not None else None
        if self.mask is not None:
            tok = nn.Parameter(torch.zeros(1, int(self.mask["channel"])))
            nn.init.trunc_normal_(tok, mean=0.0, std=0.02, a=-0.02, b=0.02)
            self.register_parameter("mtoken", tok)
        self.conditions = tuple(conditions) if conditions is not None else None
        if self.conditions is not None:   # PPT context table (:66; consumed by the PDNorm backbones, §8f-3)
            self.embedding_table = nn.Embedding(len(self.conditions), context_channels)
        # semantic branch (§8f-4, :73-118): per-class text embeddings, rendered-feature contrastive loss, PPT point loss
        self.render_semantic = bool(render_semantic)
        self.valid_index = valid_index
        self.ppt_loss_weight = float(ppt_loss_weight) if render_semantic else 0.0
        if self.render_semantic:
            emb = _class_embedding(class_embedding, template, clip_model, class_name)
            self.register_buffer("class_embedding", emb)
            self.logit_scale = nn.Parameter(torch.tensor(float(logit_scale)), requires_grad=False)   # CLIP's, frozen
            if self.ppt_loss_weight > 0:
                if ppt_criteria is None:
                    raise ValueError("PonderIndoor-v2: ppt_loss_weight > 0 needs ppt_criteria (ponder_indoor_base.py:81)")
                self.ppt_criteria = _Criteria(ppt_criteria)
                self.proj_head = nn.Linear(backbone_out_channels, emb.shape[1])

    mask_features = PonderOutdoorStep.mask_features      # the same block masking (:121-161)

    def forward(self, data_dict: Dict[str, torch.Tensor], noise: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        noise = noise or {}
        data_dict = dict(data_dict)
        if self.mask is not None:
            data_dict["feat"] = self.mask_features(data_dict["grid_coord"], data_dict["feat"], data_dict["offset"],
                                                   noise.get("mask"))
        if "condition" in data_dict and self.conditions is not None:                     # PPT context row (:164-172)
            idx = self.conditions.index(data_dict["condition"][0])
            data_dict["context"] = self.embedding_table.weight[idx:idx + 1]
        data_dict["sparse_backbone_feat"] = self.backbone(data_dict)                     # extract_feature
        cube = rayprep.to_unit_cube(data_dict)                                           # prepare_ray
        if self.render_semantic:
            cube["index2semantic"] = self._class_table(data_dict)
        ray = rayprep.ray_sample(cube, self.ray_nsample, self.bounds, pixels=noise.get("pixels"))
        cube = rayprep.grid_sample(cube, self.grid_size)                                 # prepare_volume
        cube["sparse_backbone_feat"] = data_dict["sparse_backbone_feat"]
        cube.update(ray_o=ray["ray_o"], ray_d=ray["ray_d"], rgb=ray["rgb"], depth=ray["depth"])
        cube.pop("semantic", None)                          # the (B,V,H,W) class-id maps; rays carry embeddings
        if "semantic" in ray:
            cube["semantic"] = ray["semantic"]
        out = self.forward_after_backbone(cube, noise)
        if self.ppt_loss_weight > 0:                                                     # ppt_loss (:680-691)
            feat = F.normalize(self.proj_head(data_dict["sparse_backbone_feat"].float()), dim=-1)
            logits = self.logit_scale.exp() * (feat @ self._class_table(data_dict).t())
            out["ppt_loss"] = self.ppt_criteria(logits, data_dict["segment"])
        return out

    def _class_table(self, data_dict) -> torch.Tensor:
        """Rows of `class_embedding` valid for the batch's condition (:514-523)."""
        if "condition" in data_dict and self.valid_index is not None and self.conditions is not None:
            idx = torch.as_tensor(self.valid_index[self.conditions.index(data_dict["condition"][0])],
                                  device=self.class_embedding.device)
            return self.class_embedding[idx]
        return self.class_embedding


class PonderOutdoor(PonderOutdoorStep):
    """`PonderOutdoor` with the reference's constructor (ponder_outdoor_base.py:19-91): per-condition tuples of
    scene_bbox / grid_shape / grid_size, the condition picked from data_dict["condition"][0]."""

    def __init__(self, backbone, projection, renderer, mask=None, scene_bbox=((-54.0, -54.0, -5.0, 54.0, 54.0, 3.0),),
                 grid_shape=((180, 180, 5),), grid_size=((0.6, 0.6, 1.6),), val_ray_split=8192, pool_type="mean",
                 share_volume=True, render_semantic=False, conditions=None, template=None, clip_model=None,
                 class_name=None, valid_index=None):
        if render_semantic:
            raise NotImplementedError("PonderOutdoor-v2: render_semantic=True needs CLIP text embeddings (SURVEY §8f-4)")
        nn.Module.__init__(self)
        if pool_type != "mean":
            raise NotImplementedError("pool_type: every shipped config uses 'mean'")
        nest = lambda v: tuple(tuple(x) for x in v) if isinstance(v[0], (tuple, list)) else (tuple(v),)
        self._bboxes, self._shapes, self._sizes = nest(scene_bbox), nest(grid_shape), nest(grid_size)
        self.conditions = tuple(conditions) if conditions is not None else None
        self._select(0)
        self.val_ray_split = int(val_ray_split)
        self.backbone = MODELS.build(dict(backbone))
        self.proj_net = _build_projection(projection, dict(type="SimpleConv3D-v1m1", in_channels=96, out_channels=32))
        self.renderer = RENDERERS.build(dict(renderer))
        self.mask = dict(mask) if mask is not None else None
        if self.mask is not None:
            tok = nn.Parameter(torch.zeros(1, int(self.mask["channel"])))
            nn.init.trunc_normal_(tok, mean=0.0, std=0.02, a=-0.02, b=0.02)
            self.register_parameter("mtoken", tok)

    def _select(self, i: int) -> None:
        self.scene_bbox = tuple(float(v) for v in self._bboxes[i])
        self.grid_shape = tuple(int(v) for v in self._shapes[i])
        self.grid_size = tuple(float(v) for v in self._sizes[i])

    def forward(self, data_dict, noise=None):
        if "condition" in data_dict and self.conditions is not None:
            self._select(self.conditions.index(data_dict["condition"][0]))
        return super().forward(data_dict, noise)


MODELS.register_module("PonderIndoor-v2", module=PonderIndoor)
MODELS.register_module("PonderOutdoor-v2", module=PonderOutdoor)


def install_into_reference() -> None:
    """Make the reference tree use this library: alias the third-party module names it imports and register the classes
    above in its registries under the reference's names (ponder/utils/registry.py:238-249 `register_module(force=True)`).
    Call before building a model from a reference config; requires `ponder` (the reference) on sys.path."""
    import sys

    from . import smooth_sampler as _ss
    from . import spconv as _sp
    from .spconv import pytorch as _sp_pt
    sys.modules.setdefault("spconv", _sp)
    sys.modules.setdefault("spconv.pytorch", _sp_pt)
    sys.modules.setdefault("smooth_sampler", _ss)
    from ponder.models.builder import MODELS as REF_MODELS                      # noqa: E402  (the reference)
    from ponder.models.ponder.render_utils.builder import RENDERERS as REF_RENDERERS
    for name in ("SpUNet-v1m1", "SpUNet-v1m3", "SimpleConv3D-v1m1", "UNet3D-v1m2", "PonderIndoor-v2", "PonderOutdoor-v2"):
        REF_MODELS.register_module(name=name, force=True, module=MODELS.get(name))
    REF_RENDERERS.register_module(name="NeuSModel", force=True, module=NeuSModel)
