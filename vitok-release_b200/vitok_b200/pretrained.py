"""Pretrained-weight registry and safetensors ingest (mirror of vitok/pretrained.py:1-92).

Same names, same return shape (``{'variant': ..., 'encoder': state_dict, 'decoder': state_dict}``) and the same
KeyError for unknown names, so ``model.load_state_dict({**data['encoder'], **data['decoder']})`` (README.md:50-53) and
the encoder-only / decoder-only recipes (README.md:68-82) work unchanged.  Two additions for air-gapped machines:
``local_dir`` (or $VITOK_WEIGHTS_DIR) is searched for ``<name>/encoder.safetensors`` / ``<name>/decoder.safetensors``
before the Hub is contacted, and ``save_pretrained`` writes a model in that two-file layout.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

# name -> (Hub repo, file names, variant string); identical to the reference registry (vitok/pretrained.py:8-21)
_FILES = ["encoder.safetensors", "decoder.safetensors"]
_MODELS: Dict[str, Tuple[str, List[str], str]] = {
    f"{size}-f{stride}x{ch}": (f"philippehansen/ViTok-v2-{size}-f{stride}x{ch}", list(_FILES), f"{arch}/1x{stride}x{ch}")
    for size, arch, combos in (
        ("350M", "Ld4-Ld24", ((16, 16), (16, 32), (16, 64))),
        ("5B", "Td4-T", ((16, 16), (16, 32), (16, 64), (32, 64), (32, 128), (32, 256))),
    )
    for stride, ch in combos
}
PRETRAINED_ALIASES = {name: name for name in _MODELS}

_ENCODER_PREFIXES = ("patch_embed.", "to_code.", "encoder_blocks.", "output_fn.")
_DECODER_PREFIXES = ("decoder_embed.", "to_pixels.", "decoder_blocks.")


def _check(name: str) -> Tuple[str, List[str], str]:
    if name not in _MODELS:
        raise KeyError(f"Unknown model: {name}. Available: {list(_MODELS.keys())}")
    return _MODELS[name]


def _resolve(name: str, filename: str, cache_dir: Optional[str], local_dir: Optional[str]) -> str:
    repo_id = _check(name)[0]
    for root in (local_dir, os.environ.get("VITOK_WEIGHTS_DIR")):
        if root:
            for cand in (os.path.join(root, name, filename), os.path.join(root, repo_id.split("/")[-1], filename)):
                if os.path.exists(cand):
                    return cand
    try:
        from huggingface_hub import hf_hub_download
    except ImportError as e:  # pragma: no cover
        raise ImportError("huggingface_hub is needed to download ViTok weights; pass local_dir=... for local files") from e
    return hf_hub_download(repo_id=repo_id, filename=filename, cache_dir=cache_dir)


def load_pretrained(name: str, component: Optional[str] = None, cache_dir: Optional[str] = None,
                    local_dir: Optional[str] = None) -> dict:
    """Load weights as ``{'variant', 'encoder', 'decoder'}`` (``component`` = 'encoder' | 'decoder' | None for both)."""
    _, filenames, variant = _check(name)
    from safetensors.torch import load_file
    result = {"variant": variant}
    if component != "decoder":
        result["encoder"] = load_file(_resolve(name, filenames[0], cache_dir, local_dir))
    if component != "encoder":
        result["decoder"] = load_file(_resolve(name, filenames[1], cache_dir, local_dir))
    return result


def list_pretrained() -> List[str]:
    return list(_MODELS.keys())


def get_pretrained_info(name: str) -> Tuple[str, List[str], str]:
    """(repo_id, filenames, variant) without touching the network."""
    return _check(name)


def download_pretrained(name: str, cache_dir: Optional[str] = None, local_dir: Optional[str] = None) -> List[str]:
    return [_resolve(name, f, cache_dir, local_dir) for f in _check(name)[1]]


def split_state_dict(state_dict: Dict) -> Tuple[Dict, Dict]:
    """Split a full AE state dict into the encoder / decoder halves the published checkpoints are stored as."""
    enc = {k: v for k, v in state_dict.items() if k.startswith(_ENCODER_PREFIXES)}
    dec = {k: v for k, v in state_dict.items() if k.startswith(_DECODER_PREFIXES)}
    other = set(state_dict) - set(enc) - set(dec)
    if other:
        raise KeyError(f"keys that belong to neither half: {sorted(other)[:5]}")
    return enc, dec


def save_pretrained(model, directory: str, name: str) -> List[str]:
    """Write ``<directory>/<name>/{encoder,decoder}.safetensors`` from an AE (whichever halves it has)."""
    from safetensors.torch import save_file
    enc, dec = split_state_dict({k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()})
    out_dir = os.path.join(directory, name)
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for half, fname in ((enc, _FILES[0]), (dec, _FILES[1])):
        if half:
            paths.append(os.path.join(out_dir, fname))
            save_file(half, paths[-1])
    return paths


__all__ = ["load_pretrained", "list_pretrained", "get_pretrained_info", "download_pretrained", "PRETRAINED_ALIASES",
           "save_pretrained", "split_state_dict"]
