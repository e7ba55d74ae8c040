"""Checkpoint ingestion for the drop-in module.

Mirrors the reference's loaders: `load_is_model` / `load_single_is_model` (isegm/inference/utils.py:21-46), `load_model`
(isegm/utils/serialization.py:44-69) and the two position-embedding interpolators (isegm/model/modeling/pos_embed.py:75-128).
A checkpoint is what `save_checkpoint` writes (isegm/utils/misc.py:31-33): {'state_dict': ..., 'config': net._config} with
`_config` = {'class': dotted class name, 'params': {name: {'type', 'value', 'specified'}}}.

Only the class this package implements is accepted; anything else raises (there is no fallback model).  Unpickling a
checkpoint written by the reference needs the reference package importable when its config holds reference objects
(head_params['loss_decode'] is a CrossEntropyLoss instance in the shipped configuration,
models/iSegNet/vpu_base448_cocolvis.py:34-44); pass the already loaded dict otherwise.
"""
import inspect
from pathlib import Path

import torch
import torch.nn.functional as F

from .model import REFERENCE_CLASS, VitMultiGaussianVector_ed_Model


def load_model(config, eval_ritm=False, **kwargs):
    """serialization.py:44-69: rebuild the model from a recorded constructor configuration."""
    cls_name = config["class"]
    if cls_name != REFERENCE_CLASS and cls_name.split(".")[-1] != "VitMultiGaussianVector_ed_Model":
        raise NotImplementedError("checkpoint of class %r: only %s has a B200 path" % (cls_name, REFERENCE_CLASS))
    if eval_ritm:
        raise NotImplementedError("eval_ritm (use_rgb_conv=True RITM models) is outside the B200 path")
    sig = inspect.signature(VitMultiGaussianVector_ed_Model.__init__).parameters
    args = {}
    for pname, param in config["params"].items():
        if isinstance(param, dict) and set(param) >= {"type", "value", "specified"}:
            value, specified, is_class = param["value"], param["specified"], param["type"] == "class"
        else:                                   # flat {name: value} form
            value, specified, is_class = param, True, False
        if pname not in sig:
            # arguments of the reference's base classes that do not change this path (is_model.py:10-13: norm_layer, ...)
            if specified and pname not in ("norm_layer", "binary_prev_mask", "conv_extend", "with_aux_output"):
                raise NotImplementedError("constructor argument %r of the checkpoint has no counterpart in the B200 module" % pname)
            continue
        if is_class:
            continue                            # class-valued arguments (layer factories) do not exist on this path
        if not specified and sig[pname].default == value:
            continue
        args[pname] = value
    args.update(kwargs)
    return VitMultiGaussianVector_ed_Model(**args)


def load_single_is_model(state_dict, device, eval_ritm=False, **kwargs):
    """inference/utils.py:37-46."""
    model = load_model(state_dict["config"], eval_ritm, **kwargs)
    model.load_state_dict(state_dict["state_dict"], strict=True)
    for p in model.parameters():
        p.requires_grad = False
    model.to(device)
    model.eval()
    return model


def load_is_model(checkpoint, device, eval_ritm=False, **kwargs):
    """inference/utils.py:21-34: path or loaded dict (or a list of them -> (first model, all models))."""
    if isinstance(checkpoint, (str, Path)):
        state_dict = torch.load(checkpoint, map_location="cpu", weights_only=False)
    else:
        state_dict = checkpoint
    if isinstance(state_dict, list):
        models = [load_single_is_model(x, device, eval_ritm, **kwargs) for x in state_dict]
        return models[0], models
    return load_single_is_model(state_dict, device, eval_ritm, **kwargs)


def _resample(pos_tokens, orig, new, dim):
    t = pos_tokens.reshape(-1, orig[0], orig[1], dim).permute(0, 3, 1, 2)
    t = F.interpolate(t, size=tuple(new), mode="bicubic", align_corners=False)
    return t.permute(0, 2, 3, 1).flatten(1, 2)


def interpolate_pos_embed(model, checkpoint_model):
    """pos_embed.py:75-99: resample checkpoint_model['pos_embed'] (a MAE pre-training checkpoint of another resolution) in
    place to the grid of `model` (a backbone holder: .pos_embed, .patch_embed.num_patches); extra tokens are kept."""
    if "pos_embed" not in checkpoint_model:
        return
    pe = checkpoint_model["pos_embed"]
    dim = pe.shape[-1]
    num_patches = model.patch_embed.num_patches
    extra = model.pos_embed.shape[-2] - num_patches
    orig, new = int((pe.shape[-2] - extra) ** 0.5), int(num_patches ** 0.5)
    if orig != new:
        checkpoint_model["pos_embed"] = torch.cat((pe[:, :extra], _resample(pe[:, extra:], (orig, orig), (new, new), dim)), dim=1)


def interpolate_pos_embed_inference(model, infer_img_size, device):
    """pos_embed.py:102-128, as scripts/evaluate_vpumodel.py:123-128 calls it on `net.backbone` with the zoom-in target
    size.  The CUDA forward is built for one grid (448 px / patch): an equal grid is the reference's no-op; another grid
    would need a model built for that image size (`build_model(img_size=...)`), so it raises instead of silently
    resampling parameters the kernels would not use."""
    patch = model.patch_embed.patch_size
    new = (infer_img_size[0] // patch[0], infer_img_size[1] // patch[1])
    if tuple(model.patch_embed.grid_size) != new:
        raise NotImplementedError("inference grid %s differs from the model's %s: build the model with img_size=%s and load the "
                                  "checkpoint through interpolate_pos_embed" % (new, tuple(model.patch_embed.grid_size), (infer_img_size,)))
