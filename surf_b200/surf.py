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


class SuRF(nn.Module):
    def __init__(self, confs):
        super().__init__()
        self.has_vol = False
        self.range_ratios = confs.get_list("range_ratios", default=[1.0, 0.4, 0.1, 0.01])
        self.num_stage = len(self.range_ratios)
        self.implicit_surface = ImplicitSurface(confs["implicit_surface"])
        # same attribute name as the reference (surf.py:25), so `feature_network.*` checkpoint keys load
        self.feature_network = FeatureNetwork(confs["feature_network"]) if "feature_network" in confs else None
        self.volumes = None
        self.sparse_idxes = None
        self.mask_volmes = None        # (sic) attribute name of the reference, surf.py:74
        self.matching_volume = None
        self.features = None

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
        self.has_vol = True

    def extract_features(self, imgs):
        """``self.feature_network(imgs)`` of surf.py:69: coarse -> fine list of (nv, 4, H/2^i, W/2^i) feature maps."""
        if self.feature_network is None:
            raise RuntimeError("this SuRF was built from a conf without a feature_network block")
        return self.feature_network(imgs)

    def init_volumes(self, ipts):
        raise NotImplementedError(
            "SuRF.init_volumes runs the upstream FPN + torchsparse volume construction (surf.py:65-131), which is "
            "outside the B200 hot path (SURVEY.md §8); build the volumes with the reference and pass them to "
            "set_volumes()")

    def forward(self, mode, ipts, cos_anneal_ratio=1.0, step=None):
        if not self.has_vol:
            raise NotImplementedError(
                "SuRF.forward without precomputed volumes needs build_volumes (surf.py:80-131), which is upstream of "
                "the B200 hot path (SURVEY.md §8); call set_volumes() first")
        feats = self.features
        if "view_ids" in ipts:
            view_ids = ipts["view_ids"]
            feats = [f[view_ids] for f in feats]
        # lists reversed to fine -> coarse / high-res -> low-res, exactly as surf.py:159
        return self.implicit_surface(mode, ipts, self.matching_volume, self.volumes[::-1], self.sparse_idxes[::-1],
                                     self.mask_volmes[::-1], feats[::-1], feats[::-1], cos_anneal_ratio, step)
