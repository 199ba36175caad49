"""`deinstantiate` / `swap_tensors` (drop-ins for src/tinyedm/utils.py:5-34) and the reference's checkpoint format.

The reference saves Lightning checkpoints whose `hyper_parameters` are the `deinstantiate` tree of the EDM module
(edm.py:154-157: `_target_` = "module.Class" + every constructor keyword), whose `state_dict` uses the module-tree names
this package mirrors, and whose `optimizer_states[0]["ema"]` is the positional tuple of EMA tensors in `parameters()`
order (ema.py:326-336). `instantiate` here is the small subset of hydra.utils.instantiate those trees need (hydra is
not part of this image): nested `_target_` dicts, lists -> tuples; any `tinyedm.*` target resolves to the class of the
same name in `tinyedm_b200`.
"""
from __future__ import annotations

import importlib
import inspect

import torch


def deinstantiate(obj):
    """utils.py:5-27: {"_target_": "module.Class", **constructor keywords read back from same-named attributes}."""
    class_name = f"{obj.__class__.__module__}.{obj.__class__.__name__}"
    params = {}
    for name, param in inspect.signature(obj.__init__).parameters.items():
        if name == "self":
            continue
        if hasattr(obj, name):
            attribute = getattr(obj, name)
            if isinstance(attribute, torch.nn.Module):
                params[name] = deinstantiate(attribute)
            elif isinstance(attribute, tuple):
                params[name] = list(attribute)
            else:
                params[name] = attribute
        elif param.default is not inspect.Parameter.empty:
            params[name] = param.default
    return {"_target_": class_name, **params}


def swap_tensors(tensor1, tensor2):
    """utils.py:30-34 (in place: the EMA optimiser swaps `param.data` with the averaged copy)."""
    tmp = torch.empty_like(tensor1)
    tmp.copy_(tensor1)
    tensor1.copy_(tensor2)
    tensor2.copy_(tmp)
    # `param.data` writes do not bump `param._version`: tell the weight banks their cached normalised weights are stale
    from .engine import bump_weights_epoch
    bump_weights_epoch()


def _resolve(target: str):
    module, _, name = target.rpartition(".")
    if module.split(".")[0] in ("tinyedm", "tinyedm_b200"):
        import tinyedm_b200
        if hasattr(tinyedm_b200, name):
            return getattr(tinyedm_b200, name)
        module = module.replace("tinyedm", "tinyedm_b200", 1) if module.startswith("tinyedm.") or module == "tinyedm" else module
    return getattr(importlib.import_module(module), name)


def instantiate(cfg):
    """Builds the object tree of a `deinstantiate` dict (or of the `model:` node of the reference's YAML configs)."""
    if isinstance(cfg, dict):
        if "_target_" in cfg:
            cls = _resolve(cfg["_target_"])
            kwargs = {k: instantiate(v) for k, v in cfg.items() if k not in ("_target_", "_partial_", "_convert_")}
            accepted = inspect.signature(cls.__init__).parameters
            if not any(p.kind is inspect.Parameter.VAR_KEYWORD for p in accepted.values()):
                kwargs = {k: v for k, v in kwargs.items() if k in accepted}
            return cls(**kwargs)
        return {k: instantiate(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return tuple(instantiate(v) for v in cfg)
    return cfg


def find_ema_weights(checkpoint: dict):
    """edm.py:196-203."""
    try:
        return checkpoint["optimizer_states"][0]["ema"]
    except (KeyError, IndexError, TypeError):
        raise ValueError("EMA weights not found in the checkpoint.")


def load_reference_checkpoint(checkpoint, *, load_ema: bool = False, map_location=None):
    """`checkpoint`: path or already-loaded dict in the reference's format. Returns a `tinyedm_b200.EDM` carrying the
    checkpoint's weights (or, with `load_ema`, the EMA weights: edm.py:172-182)."""
    if not isinstance(checkpoint, dict):
        checkpoint = torch.load(checkpoint, map_location=map_location, weights_only=False)
    model = instantiate(dict(checkpoint["hyper_parameters"]))
    state = checkpoint.get("state_dict") or {}
    if state:
        missing, unexpected = model.load_state_dict(state, strict=False)
        real_missing = [k for k in missing if not k.startswith(("train_mse", "val_mse"))]
        if real_missing or unexpected:
            raise RuntimeError(f"checkpoint / module mismatch: missing {real_missing}, unexpected {list(unexpected)}")
    if load_ema:
        ema = find_ema_weights(checkpoint)
        params = list(model.parameters())
        if len(ema) != len(params):
            raise RuntimeError(f"checkpoint holds {len(ema)} EMA tensors for {len(params)} parameters")
        with torch.no_grad():
            for p, e in zip(params, ema):
                p.copy_(e.to(p.device, p.dtype))
    return model
