"""``--yaml-cfg``: experimental YAML alternative to the command line (reference
``training/yaml_arguments.py``).  Nested sections (``language_model:``, ``model_parallel:``,
``training:`` ...) are flattened onto the ordinary argument namespace, unknown keys are rejected,
then the usual ``validate_args`` runs -- so YAML and CLI configs are interchangeable."""
from __future__ import annotations

import os
import re
from types import SimpleNamespace

import yaml

_ENV = re.compile(r"\$\{([^}^{]+)\}")


def _env_constructor(loader, node):
    value = loader.construct_scalar(node)
    for g in _ENV.findall(value):
        value = value.replace(f"${{{g}}}", os.environ.get(g, ""))
    return value


class _Loader(yaml.SafeLoader):
    pass


_Loader.add_implicit_resolver("!env", _ENV, None)
_Loader.add_constructor("!env", _env_constructor)


def _flatten(d, out):
    for k, v in d.items():
        if isinstance(v, dict) and not k.endswith("_kwargs"):
            _flatten(v, out)
        else:
            if k in out:
                raise ValueError(f"duplicate key {k!r} in yaml config")
            out[k] = v
    return out


def load_yaml(path: str) -> SimpleNamespace:
    """Returns the raw nested config as namespaces (sections stay addressable: ``cfg.language_model``)."""
    with open(path) as f:
        data = yaml.load(f, Loader=_Loader)

    def ns(x):
        return SimpleNamespace(**{k: ns(v) for k, v in x.items()}) if isinstance(x, dict) else x

    return ns(data)


def args_from_yaml(path: str, defaults):
    """Overlay the flattened YAML onto an argparse namespace of defaults."""
    with open(path) as f:
        flat = _flatten(yaml.load(f, Loader=_Loader) or {}, {})
    known = vars(defaults)
    for k, v in flat.items():
        key = k.replace("-", "_")
        if key not in known:
            raise ValueError(f"unknown option {k!r} in {path}")
        setattr(defaults, key, v)
    defaults.yaml_cfg = path
    return defaults
