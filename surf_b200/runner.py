"""Runner.validate-shaped consumer of the drop-in (mirror of runner.py:198-296).

The reference's ``Runner`` owns datasets, optimiser, tensorboard and DDP set-up (runner.py:27-104) — control plane,
out of scope.  What the hot path must satisfy is the *consumer contract* of ``validate`` (and of the validation block
of ``finetune``, runner.py:351-395): call ``model("val", inputs, cos_anneal_ratio=...)``, read ``img_fine``,
``normal_img``, ``color_fine``, ``sdf_depth``, ``render_depth``, ``vertices``, ``triangles`` (+ ``depth_stage0`` when the
upstream matching field produced it), write the same artefacts under ``base_exp_dir`` with the same file names, and
return the same scalars.  ``trimesh`` / ``matplotlib`` are not in this image: the mesh is written by
``surf_b200.mesh.write_ply`` and the depth maps are colour-mapped with a piecewise-linear fit of *magma*.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable

import numpy as np
import torch
import torch.nn.functional as F

from . import mesh as _mesh

# anchor colours of matplotlib's "magma" at t = 0, 1/8, ..., 1 (piecewise-linear stand-in for cm.magma, host IO only)
_MAGMA = np.array([[0.001, 0.000, 0.014], [0.110, 0.066, 0.298], [0.316, 0.072, 0.485], [0.513, 0.148, 0.508],
                   [0.717, 0.215, 0.475], [0.904, 0.319, 0.388], [0.987, 0.535, 0.382], [0.996, 0.769, 0.534],
                   [0.987, 0.991, 0.750]])


def colormap_depth(depth: np.ndarray, vmin: float = 0.0, vmax: float = 3.0) -> np.ndarray:
    """runner.py:400-413 (Normalize(0, 3) + magma) -> (h, w, 3) uint8."""
    t = np.clip((np.asarray(depth, dtype=np.float64) - vmin) / (vmax - vmin), 0.0, 1.0) * (len(_MAGMA) - 1)
    i0 = np.minimum(t.astype(np.int64), len(_MAGMA) - 2)
    f = (t - i0)[..., None]
    return ((_MAGMA[i0] * (1 - f) + _MAGMA[i0 + 1] * f) * 255).astype(np.uint8)


def save_depth(depth: np.ndarray, file_path: str) -> None:
    from PIL import Image
    Image.fromarray(colormap_depth(depth)).save(file_path)


def get_cos_anneal_ratio(step: float, anneal_end: float = 0.0) -> float:
    """runner.py:415-419."""
    return 1.0 if anneal_end == 0.0 else float(np.min([1.0, step / anneal_end]))


def _to_device(v, device):
    return v.to(device) if isinstance(v, torch.Tensor) else v


def export_mesh(vertices: np.ndarray, triangles: np.ndarray, scale_mat, path: str) -> None:
    """``trimesh.Trimesh(v, f).apply_transform(scale_mat).export(path)`` (runner.py:236-243) without trimesh."""
    m = np.asarray(scale_mat.detach().cpu().numpy() if isinstance(scale_mat, torch.Tensor) else scale_mat, dtype=np.float64)
    m = m.reshape(4, 4)
    v = np.asarray(vertices, dtype=np.float64) @ m[:3, :3].T + m[:3, 3][None, :]
    _mesh.write_ply(path, v, triangles)


@torch.no_grad()
def validate(model, val_loader: Iterable[Dict], base_exp_dir: str, epoch: int = 0, anneal_end: float = 0.0,
             device="cuda:0", tag: str = "epoch", clean_mesh: bool = False, **val_kwargs) -> Dict[str, float]:
    """One validation pass: per item a rendered image, normal map, two depth maps (.png + .npy) and the mesh, exactly
    the files runner.py:243-262 writes; returns the averaged scalars of runner.py:264-286.  ``tag='step'`` gives the
    file names of the finetune validation block (runner.py:377-388).  ``val_kwargs`` are forwarded to
    ``ImplicitSurface.validate`` through ``model.val_options`` when the model supports it (e.g. mesh_resolution).
    ``clean_mesh`` = the reference's ``--clean_mesh`` switch (runner.py:233-234): the mesh is cleaned against
    ``inputs["masks"]`` on the GPU (surf_b200.clean_mesh) before it is transformed and written."""
    from PIL import Image
    model.eval()
    items = list(val_loader)
    sums: Dict[str, float] = {}
    if val_kwargs:
        target = getattr(model, "implicit_surface", model)
        target.val_options = dict(val_kwargs)
    try:
        for batch, inputs in enumerate(items):
            inputs = {k: _to_device(v, device) for k, v in inputs.items()}
            ratio = get_cos_anneal_ratio(epoch + batch / len(items), anneal_end)
            outputs = model("val", inputs, cos_anneal_ratio=ratio)
            file_name, scene = inputs.get("file_name", str(batch)), inputs.get("scene", "scene")
            color_fine, sdf_depth, render_depth = outputs["color_fine"], outputs["sdf_depth"], outputs["render_depth"]
            sfx = "_%s%s" % (tag, epoch)
            for sub in ("meshes", "val_img", "val_normal", "val_sdf_depth", "val_render_depth"):
                os.makedirs(os.path.join(base_exp_dir, sub), exist_ok=True)
            if "vertices" in outputs:
                if clean_mesh:          # runner.py:233-234
                    from .clean_mesh import clean_mesh as _clean
                    outputs["vertices"], outputs["triangles"] = _clean(outputs["vertices"], outputs["triangles"],
                                                                       inputs["masks"], inputs["intrs"], inputs["c2ws"],
                                                                       device=device)
                export_mesh(outputs["vertices"], outputs["triangles"], inputs.get("scale_mat", np.eye(4)),
                            os.path.join(base_exp_dir, "meshes", "%s%s.ply" % (scene, sfx)))
            Image.fromarray(outputs["img_fine"].astype(np.uint8)).save(os.path.join(base_exp_dir, "val_img", file_name + sfx + ".png"))
            Image.fromarray(outputs["normal_img"].astype(np.uint8)).save(os.path.join(base_exp_dir, "val_normal", file_name + sfx + ".png"))
            save_depth(render_depth, os.path.join(base_exp_dir, "val_render_depth", file_name + sfx + ".png"))
            save_depth(sdf_depth, os.path.join(base_exp_dir, "val_sdf_depth", file_name + sfx + ".png"))
            np.save(os.path.join(base_exp_dir, "val_render_depth", file_name + sfx + ".npy"), render_depth)
            np.save(os.path.join(base_exp_dir, "val_sdf_depth", file_name + sfx + ".npy"), sdf_depth)
            scalars: Dict[str, float] = {}
            if "color" in inputs:
                gt = inputs["color"].cpu()
                scalars["psnr"] = float(20.0 * torch.log10(1.0 / (((color_fine - gt) ** 2).mean()).sqrt()))
                scalars["color_loss"] = float(F.l1_loss(color_fine, gt))
            auxi = outputs.get("depth_stage0")
            if auxi is not None:        # produced by the upstream matching field (surf.py:161), not by the hot path
                auxi = auxi.detach().cpu().numpy()
                os.makedirs(os.path.join(base_exp_dir, "val_auxi_depth"), exist_ok=True)
                save_depth(auxi, os.path.join(base_exp_dir, "val_auxi_depth", file_name + sfx + ".png"))
                np.save(os.path.join(base_exp_dir, "val_auxi_depth", file_name + sfx + ".npy"), auxi)
            if "depth_ref" in inputs:
                depth_ref = inputs["depth_ref"].cpu().numpy()
                skip = depth_ref.shape[0] // render_depth.shape[0]
                depth_ref = depth_ref[::skip, ::skip]
                mask_ref = depth_ref > 0
                scalars["render_depth_loss"] = float((np.abs(render_depth - depth_ref) * mask_ref).sum() / (mask_ref.sum() + 1e-8))
                m2 = mask_ref * (sdf_depth > 0)
                scalars["sdf_depth_loss"] = float((np.abs(sdf_depth - depth_ref) * m2).sum() / (m2.sum() + 1e-8))
                if auxi is not None:
                    a = auxi[::skip, ::skip]
                    scalars["auxi_depth_loss"] = float((np.abs(a - depth_ref) * mask_ref).sum() / (mask_ref.sum() + 1e-8))
            for k, v in scalars.items():
                sums[k] = sums.get(k, 0.0) + v
    finally:
        if val_kwargs:
            target.val_options = {}
    return {k: v / max(1, len(items)) for k, v in sums.items()}
