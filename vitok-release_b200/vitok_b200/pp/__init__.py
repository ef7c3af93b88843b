"""NaFlex pre/post-processing (mirror of vitok/pp/__init__.py:23-27)."""
from .ops import OPS, pack_images, patchify_batch, patchify_packed, unpack, unpatchify
from .registry import build_transform, parse_op
from .io import postprocess, postprocess_images, preprocess, preprocess_images

__all__ = ["build_transform", "parse_op", "OPS", "preprocess", "postprocess", "preprocess_images", "postprocess_images",
           "unpatchify", "unpack", "patchify_batch", "pack_images", "patchify_packed"]
