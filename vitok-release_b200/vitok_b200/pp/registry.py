"""String DSL -> transform pipeline, e.g. ``"to_tensor|normalize(minus_one_to_one)|patchify(16, 256)"``.

Host-side mirror of vitok/pp/registry.py (same grammar: ``name`` or ``name(args, kw=val)`` segments
joined by ``|``; bare identifiers inside the parentheses are strings).  Same errors: ValueError for
empty / malformed segments, KeyError for unknown ops (registry.py:26-31, 90-92).
"""
from __future__ import annotations

import ast
import re
from typing import Any, Callable, Dict, List, Tuple

from .ops import OPS

_SEGMENT = re.compile(r"^(\w+)(?:\((.*)\))?$", re.DOTALL)


def _literal(node: ast.AST) -> Any:
    if isinstance(node, ast.Name):  # normalize(minus_one_to_one): unquoted word = string
        return node.id
    return ast.literal_eval(ast.unparse(node))


def parse_op(op_str: str) -> Tuple[str, Tuple[Any, ...], Dict[str, Any]]:
    text = op_str.strip()
    if not text:
        raise ValueError("Empty op string")
    m = _SEGMENT.match(text)
    if m is None:
        raise ValueError(f"Invalid op syntax: '{text}'")
    name, inner = m.group(1), m.group(2)
    if inner is None or not inner.strip():
        return name, (), {}
    try:
        call = ast.parse(f"_({inner})", mode="eval").body
    except SyntaxError as exc:
        raise ValueError(f"Invalid arguments in '{text}': {exc}")
    return name, tuple(_literal(a) for a in call.args), {k.arg: _literal(k.value) for k in call.keywords}


def parse_pipeline(pp_string: str) -> List[Tuple[str, Tuple[Any, ...], Dict[str, Any]]]:
    steps = []
    for seg in (pp_string or "").split("|"):
        if not seg.strip():
            continue
        name, args, kwargs = parse_op(seg)
        if name not in OPS:
            raise KeyError(f"Unknown op: '{name}'. Available: {', '.join(sorted(OPS))}")
        steps.append((name, args, kwargs))
    return steps


def build_transform(pp_string: str) -> Callable:
    fns = [OPS[name](*args, **kwargs) for name, args, kwargs in parse_pipeline(pp_string)]

    def run(x):
        for fn in fns:
            x = fn(x)
        return x

    return run


__all__ = ["build_transform", "parse_op", "parse_pipeline", "OPS"]
