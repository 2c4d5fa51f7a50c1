"""Config loading for the drop-in: the reference's YAMLs load unchanged.

Mirrors what the reference gets from OmegaConf (absent here; PyYAML is present):
  * ``${from_file:rel/path.yaml}`` and ``${negation:...}`` resolvers   (reference main.py:72-76)
  * plain ``${a.b.c}`` interpolation                                    (OmegaConf built-in)
  * ``key=value`` dot-list overrides                                    (reference main.py:49)
  * ``instantiate_from_config({"target", "params"})``                   (reference utils/utils.py:11-22)

``target`` strings that name reference classes on the hot path are mapped onto this package's
B200 implementations (the plugin boundary); anything else is imported by dotted path
exactly as the reference does.
"""
from __future__ import annotations

import importlib
import os
import re
from typing import Any, Dict, List, Optional

import yaml

# reference dotted path -> implementation in this package
TARGET_MAP = {
    "models.modules.sampler.llama.Transformer": "vaura_b200.sampler.Transformer",
    "models.modules.dac.model.DacModelWrapper": "vaura_b200.codec.DacModelWrapper",
    "models.modules.misc.codebook_patterns.DelayedPatternProvider": "vaura_b200.patterns.DelayedPatternProvider",
    "models.modules.feature_extractors.avclip.motionformer.MotionFormer": "vaura_b200.features.MotionFormer",
    "models.vaura_model.VAURAModel": "vaura_b200.model.VAURAModel",
}

_INTERP = re.compile(r"\$\{([^${}]+)\}")


class _Loader(yaml.SafeLoader):
    pass


# YAML 1.1 does not read "1e-5" as a float; OmegaConf does.  Add the same implicit resolver.
_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                    |\.[0-9_]+(?:[eE][-+][0-9]+)?
                    |[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$""", re.X),
    list("-+0123456789."),
)


def load_yaml(path: str) -> Any:
    with open(path) as f:
        return yaml.load(f, Loader=_Loader)


def get_obj_from_str(string: str):
    string = TARGET_MAP.get(string, string)
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config: Dict[str, Any]):
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**(config.get("params") or dict()))


def _select(root: Dict[str, Any], dotted: str):
    node = root
    for part in dotted.split("."):
        node = node[int(part)] if isinstance(node, list) else node[part]
    return node


def _coerce(text: str):
    return yaml.load(text, Loader=_Loader)


def _resolve_str(value: str, root, base_dir: str, depth: int = 0):
    if depth > 32:
        raise ValueError(f"interpolation too deep in {value!r}")
    while True:
        m = _INTERP.search(value)
        if m is None:
            return value
        expr = m.group(1)
        if expr.startswith("from_file:"):
            path = expr[len("from_file:"):].strip()
            path = path if os.path.isabs(path) else os.path.join(base_dir, path)
            res = resolve(load_yaml(path), base_dir=base_dir)
        elif expr.startswith("negation:"):
            arg = expr[len("negation:"):].strip()
            arg = _coerce(arg) if isinstance(arg, str) else arg
            res = not bool(arg)
        else:
            res = _select(root, expr.strip())
            if isinstance(res, str):
                res = _resolve_str(res, root, base_dir, depth + 1)
        if m.span() == (0, len(value)):
            return res
        value = value[: m.start()] + str(res) + value[m.end():]


def resolve(cfg: Any, base_dir: str = ".", root: Optional[Any] = None) -> Any:
    """Resolve every interpolation in place and return ``cfg``.  ``???`` (OmegaConf's
    mandatory-missing marker) is left as is, as OmegaConf.resolve does."""
    root = cfg if root is None else root
    if isinstance(cfg, dict):
        for k in list(cfg.keys()):
            cfg[k] = resolve(cfg[k], base_dir, root)
    elif isinstance(cfg, list):
        for i in range(len(cfg)):
            cfg[i] = resolve(cfg[i], base_dir, root)
    elif isinstance(cfg, str) and "${" in cfg:
        return _resolve_str(cfg, root, base_dir)
    return cfg


def merge(base: Any, override: Any) -> Any:
    """OmegaConf.merge semantics for dict trees: recursive on dicts, replace otherwise."""
    if isinstance(base, dict) and isinstance(override, dict):
        out = dict(base)
        for k, v in override.items():
            out[k] = merge(base[k], v) if k in base else v
        return out
    return override


def from_dotlist(items: List[str]) -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    for item in items:
        key, _, val = item.partition("=")
        node = out
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = _coerce(val) if val != "" else None
    return out


def load_config(path: str, overrides: Optional[List[str]] = None, base_dir: Optional[str] = None) -> Dict[str, Any]:
    """Load a reference config file (e.g. configs/generate_vgg.yaml or an experiment YAML).

    ``base_dir`` is the directory relative ``from_file`` paths are resolved against — the
    reference resolves them against the process CWD (its repo root); pass that root here."""
    cfg = load_yaml(path)
    if overrides:
        cfg = merge(cfg, from_dotlist(overrides))
    return resolve(cfg, base_dir=base_dir or os.getcwd())


# ---- checkpoint pickers of the generate driver (reference utils/utils.py:25-45) -----------------------------------
def get_latest_file(path, pattern: str = "*"):
    """Newest file (ctime) under ``path`` matching ``pattern``."""
    from pathlib import Path

    files = list(Path(path).glob(pattern))
    if not files:
        raise FileNotFoundError(f"No files found in {path} with pattern {pattern}")
    return max(files, key=lambda x: x.stat().st_ctime)


def get_file_with_best_val_loss(path, pattern: str = "*.ckpt"):
    """The checkpoint whose file name carries the lowest ``val_loss=<float>`` (Lightning's ModelCheckpoint naming);
    a directory with a single checkpoint returns it whatever its name."""
    from pathlib import Path

    ckpts = sorted(Path(path).glob(pattern))
    assert len(ckpts) > 0, f"No files found in {path} with pattern {pattern}"
    if len(ckpts) == 1:
        return ckpts[0]
    best, best_loss = None, float("inf")
    for f in ckpts:
        m = re.search(r"val_loss=([0-9]+(?:\.[0-9]+)?)", f.name)
        if m and float(m.group(1)) < best_loss:
            best, best_loss = f, float(m.group(1))
    if best is None:
        raise FileNotFoundError(f"no checkpoint name under {path} carries val_loss=<float>")
    return best
