"""SuRF — the model top for the render hot path (mirror of models/surf.py:15-163).

Scope: this class owns ``implicit_surface`` (same attribute name, so
``implicit_surface.*`` state_dict keys match the reference checkpoint) and dispatches ``forward`` to it
with the scene lists reversed to renderer order (surf.py:159).  Volume construction
(``build_volumes``, surf.py:80-131) is upstream of the hot path: its pieces exist here as drop-in modules —
``feature_network`` (the 2-D pyramid, ``modules/feature_network.py``), ``modules/volume.py`` and
``modules/matching_field.py`` — except the torchsparse cost-volume regularisation (``reg_network``), so scenes still
enter through ``set_volumes`` (the ``has_vol`` branch of the reference, surf.py:149-157) in the reference's own
coarse->fine layout.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .modules.feature_network import FeatureNetwork
from .modules.implicit_surface import ImplicitSurface
from .modules.matching_field import MatchingField
from .modules.volume import Volume


class SuRF(nn.Module):
    def __init__(self, confs):
        super().__init__()
        self.has_vol = False
        self.range_ratios = confs.get_list("range_ratios", default=[1.0, 0.4, 0.1, 0.01])
        self.num_stage = len(self.range_ratios)
        self.implicit_surface = ImplicitSurface(confs["implicit_surface"])
        # same attribute name as the reference (surf.py:25), so `feature_network.*` checkpoint keys load
        self.feature_network = FeatureNetwork(confs["feature_network"]) if "feature_network" in confs else None
        # the frozen copy the reference renders the matching features with (surf.py:30-32, 141-148)
        self.match_feature_network = FeatureNetwork(confs["feature_network"]) if "feature_network" in confs else None
        if self.match_feature_network is not None:
            for p in self.match_feature_network.parameters():
                p.requires_grad = False
        self.volume = Volume(confs["volume"]) if "volume" in confs else None
        self.matching_field = MatchingField(confs["matching_field"]) if "matching_field" in confs else None
        # The cost-volume regularisation network is the one piece of build_volumes that is not built here (torchsparse,
        # SURVEY §8f F4).  Plug one in: a callable (feats (n, c) fp32, coords (n, 4) int32 [batch, x, y, z], stage) ->
        # (out_feats (n, 8), mid_feats (n, d_base)); `TorchsparseRegNetwork` adapts the reference's module.
        self.reg_network = None
        self.volumes = None
        self.sparse_idxes = None
        self.mask_volmes = None        # (sic) attribute name of the reference, surf.py:74
        self.matching_volume = None
        self.features = None
        self.prepared = None           # init_volumes(compact=True): the scene in the prepared layout, no reference tensors

    def get_optim_params(self, lr_conf):
        return [{"params": list(self.implicit_surface.parameters()), "lr": lr_conf["mlp_lr"]}]

    def set_volumes(self, volumes_all, sparse_idx_all, mask_volumes_all, matching_volume, features):
        """Install scene tensors in the layout ``build_volumes`` returns (coarse -> fine lists;
        ``features`` coarse -> fine as ``feature_network`` emits them).  Equivalent of ``init_volumes``
        (surf.py:65-78) with the upstream networks factored out."""
        self.volumes = list(volumes_all)
        self.sparse_idxes = list(sparse_idx_all)
        self.mask_volmes = list(mask_volumes_all)
        self.matching_volume = matching_volume
        self.features = list(features)
        self.prepared = None
        self.has_vol = True

    def extract_features(self, imgs):
        """``self.feature_network(imgs)`` of surf.py:69: coarse -> fine list of (nv, 4, H/2^i, W/2^i) feature maps."""
        if self.feature_network is None:
            raise RuntimeError("this SuRF was built from a conf without a feature_network block")
        return self.feature_network(imgs)

    def _upstream_ready(self):
        missing = [n for n in ("feature_network", "volume", "matching_field", "reg_network") if getattr(self, n) is None]
        if missing:
            raise NotImplementedError(
                "SuRF.build_volumes needs %s; the torchsparse cost-volume regularisation network (reg_network.py) is not "
                "part of surf_b200 — assign `model.reg_network` (e.g. TorchsparseRegNetwork(reference_module)) or build "
                "the volumes with the reference and pass them to set_volumes()" % ", ".join(missing))

    @torch.no_grad()
    def build_volumes(self, ipts, features, perturb=False, compact=False):
        """The coarse-to-fine volume construction of surf.py:80-131 on the kernels of Volume / MatchingField, with the
        regularisation network supplied by the caller.  Same returns: (outputs with depth_stage{s} / depth_src_stage{s},
        volumes_all, sparse_idx_all, mask_volumes_all, matching_volume), lists coarse -> fine.
        ``compact=True``: the int64 index tables and fp32 mask volumes of the reference layout (6 GB at 704^3, 45 GB at
        1408^3) are never built; returns (outputs, PreparedScene) with the scene in the renderer's own layout (int32
        index, 1-bit masks) straight from the per-level voxel lists."""
        self._upstream_ready()
        intrs, c2ws = ipts["intrs"], ipts["c2ws"]
        base_range = (ipts["far"] - ipts["near"]).squeeze()
        vol = self.volume
        volumes_all, sparse_idx_all, mask_volumes_all = [], [], []
        depths, matching_volume, coords, carried = None, None, None, None
        coords_all, logits_all, dims_all = [], [], []
        outputs = {}
        for s in range(self.num_stage):
            if s == 0:
                coords = vol.init_coords().to(intrs)
                up_feats = None
            else:
                coords, up_feats = vol.up_sample(coords, carried)
                coords, up_feats = vol.depth_filtering(depths, coords, up_feats, intrs, c2ws,
                                                       base_range * self.range_ratios[s])
            feats, in_frustum = vol.back_proj_multiscale(features, coords, intrs, c2ws, s)
            feats, coords = feats[in_frustum], coords[in_frustum]
            if up_feats is not None:
                feats = torch.cat([feats, up_feats[in_frustum]], dim=1)
            batched = torch.cat([torch.zeros_like(coords[:, :1]), coords], dim=1).to(torch.int32)     # batch first (:112)
            out_feats, carried = self.reg_network(feats, batched, s)
            matching_volume, mask_volume = vol.sparse2dense(out_feats[:, :1], coords, matching_volume)
            volumes_all.append(out_feats[:, 1:].contiguous())
            if compact:
                del mask_volume
                coords_all.append(coords.clone())
                logits_all.append(out_feats[:, :1].contiguous())
                dims_all.append(int(vol.volume_dim[0]))
            else:
                sparse_idx_all.append(vol.get_index(coords))
                mask_volumes_all.append(mask_volume)
            depths, _ = self.matching_field(ipts, matching_volume, s, self.range_ratios, depths, perturb=perturb)
            outputs["depth_stage%d" % s] = depths[0]
            outputs["depth_src_stage%d" % s] = depths[ipts["src_idx"]] if "src_idx" in ipts else depths[0]
        if compact:
            return outputs, Volume.to_prepared_scene(coords_all, volumes_all, logits_all, dims_all)
        return outputs, volumes_all, sparse_idx_all, mask_volumes_all, matching_volume

    def init_volumes(self, ipts, compact=False):
        """surf.py:65-78: features + volumes of a scene from its images; afterwards forward() renders.
        ``compact=True`` keeps the scene only in the renderer's prepared layout (see build_volumes)."""
        features = self.extract_features(ipts["imgs"])               # coarse to fine
        if compact:
            _, scene = self.build_volumes(ipts, features, False, compact=True)
            self.volumes = self.sparse_idxes = self.mask_volmes = self.matching_volume = None
            self.features = list(features)
            self.prepared = scene
            self.has_vol = True
            return
        _, volumes, sparse_idxes, mask_volumes, matching_volume = self.build_volumes(ipts, features, False)
        self.set_volumes(volumes, sparse_idxes, mask_volumes, matching_volume, features)

    def forward(self, mode, ipts, cos_anneal_ratio=1.0, step=None):
        outputs = {}
        if not self.has_vol:
            # the generalisable path (surf.py:137-148): pyramid + volumes are rebuilt from the images on every call
            self._upstream_ready()
            imgs = ipts["imgs"]
            feats = self.feature_network(imgs)
            mf_outputs, volumes, sparse_idxes, mask_volumes, matching_volume = self.build_volumes(ipts, feats,
                                                                                                 perturb=(mode == "train"))
            outputs.update(mf_outputs)
            if step is not None and step % 2 == 0:
                self.match_feature_network.load_state_dict(self.feature_network.state_dict(), strict=True)
            match_feats = self.match_feature_network(imgs)
        else:
            feats = self.features
            if self.prepared is not None:        # compact scene: only the views change per call
                if "view_ids" in ipts:
                    feats = [f[ipts["view_ids"]] for f in feats]
                outputs.update(self.implicit_surface(mode, ipts, self.prepared, None, None, None, feats[::-1], feats[::-1],
                                                     cos_anneal_ratio, step))
                return outputs
            volumes, sparse_idxes, mask_volumes = self.volumes, self.sparse_idxes, self.mask_volmes
            matching_volume = self.matching_volume
            if "view_ids" in ipts:
                view_ids = ipts["view_ids"]
                feats = [f[view_ids] for f in feats]
            match_feats = feats
        # lists reversed to fine -> coarse / high-res -> low-res, exactly as surf.py:159
        outputs.update(self.implicit_surface(mode, ipts, matching_volume, volumes[::-1], sparse_idxes[::-1],
                                             mask_volumes[::-1], feats[::-1], match_feats[::-1], cos_anneal_ratio, step))
        return outputs


class TorchsparseRegNetwork:
    """Adapter for the reference's ``SparseCostRegNetList`` (models/modules/reg_network.py:87-106; needs torchsparse):
    ``model.reg_network = TorchsparseRegNetwork(reference_reg_network)``."""

    def __init__(self, module):
        from torchsparse.tensor import SparseTensor        # raises where torchsparse is not installed
        self._sparse_tensor, self._module = SparseTensor, module

    def __call__(self, feats, coords, stage):
        return self._module(self._sparse_tensor(feats=feats, coords=coords), stage)
