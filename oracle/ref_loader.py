"""Loads the byte-compiled reference modules from `oracle/_ref/` (see oracle/build_ref.py) — TEST / BASELINE
INFRASTRUCTURE ONLY; never imported by the product package.

`load()` returns an object with `.networks` / `.solvers` (the reference's own `src/tinyedm/networks.py`, `solvers.py`)
or None when the bytecode is absent or was built by another Python. `reference_edm_parts(cfg)` builds the reference's
`Denoiser` / `Embedding` for one of the oracle's config dicts and `load_params` copies an oracle parameter dict into them
(the state_dict names are the reference's by construction).
"""
from __future__ import annotations

import marshal
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_cache: dict = {}


def _load_one(name: str):
    path = os.path.join(REF_DIR, name + ".code")
    if not os.path.exists(path):
        return None
    with open(path, "rb") as fh:
        code = marshal.loads(fh.read())
    mod = types.ModuleType(f"tinyedm_ref_{name}")
    mod.__file__ = path
    sys.modules[mod.__name__] = mod
    exec(code, mod.__dict__)
    return mod


def load():
    if "ref" not in _cache:
        ref = None
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                nets, solv = _load_one("networks"), _load_one("solvers")
            if nets is not None and solv is not None:
                ref = types.SimpleNamespace(networks=nets, solvers=solv)
        except Exception as e:  # noqa: BLE001 - stale bytecode (other interpreter): behave as absent
            ref = None
            _cache["error"] = f"{type(e).__name__}: {e}"
        _cache["ref"] = ref
    return _cache["ref"]


def load_error() -> str | None:
    """Why `load()` returned None (None when the bytecode is simply absent)."""
    load()
    return _cache.get("error")


def reference_edm_parts(cfg: dict):
    """(Denoiser, Embedding) of the REFERENCE for an oracle config dict (oracle.edm2_oracle.CIFAR10 / MNIST / IMAGENET / SMALL)."""
    ref = load()
    if ref is None:
        raise RuntimeError("oracle/_ref is not built (python oracle/build_ref.py in the build container)")
    s, e = cfg["denoiser"], cfg["embedding"]
    den = ref.networks.Denoiser(
        in_channels=s.in_channels, out_channels=s.out_channels, encoder_block_types=s.encoder_block_types,
        decoder_block_types=s.decoder_block_types, encoder_out_channels=s.encoder_out_channels,
        decoder_out_channels=s.decoder_out_channels, skip_connections=s.skip_connections, dropout_rate=s.dropout_rate,
        sigma_data=s.sigma_data, encoder_add_factor=s.encoder_add_factor, decoder_add_factor=s.decoder_add_factor,
        embedding_dim=s.embedding_dim, num_heads=s.num_heads)
    emb = ref.networks.Embedding(e.fourier_dim, e.embedding_dim, e.num_classes, e.add_factor)
    return den, emb


def load_params(module, params: dict) -> None:
    module.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
