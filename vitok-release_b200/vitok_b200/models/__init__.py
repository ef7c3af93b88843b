from .ae import AE, Model, decode_variant

__all__ = ["AE", "Model", "decode_variant"]
