"""Minimal config tree for the hot path.

The reference drives everything from an OmegaConf structured config (ganslate/configs/utils.py:10-90): a YAML
tree where every `_target_: pkg.Class` node is completed with the defaults of the sibling dataclass
`ClassConfig` found in the same module (configs/utils.py:55-61).  OmegaConf is not in this image, so `Conf`
offers the part of DictConfig the hot path reads: attribute + item access, `in`, iteration, `dict(node)`.
When OmegaConf is installed a DictConfig can be passed to the recipes directly -- they only read attributes.
"""
import dataclasses
from typing import Any, Mapping

from ganslate_b200.utils.io import import_attr


class Conf(Mapping):

    def __init__(self, data=None):
        object.__setattr__(self, "_d", {})
        for k, v in (data or {}).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, Conf):
            return v
        if dataclasses.is_dataclass(v) and not isinstance(v, type):
            v = dataclasses.asdict(v)
        if isinstance(v, Mapping):
            return Conf(v)
        if isinstance(v, list):
            return tuple(v)
        return v

    def __getattr__(self, k):
        try:
            return self._d[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self._d[k] = self._wrap(v)

    def __getitem__(self, k):
        return self._d[k]

    def __setitem__(self, k, v):
        self._d[k] = self._wrap(v)

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __contains__(self, k):
        return k in self._d

    def get(self, k, default=None):
        return self._d.get(k, default)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Conf) else v) for k, v in self._d.items()}

    def merge(self, other: Mapping):
        for k, v in other.items():
            if isinstance(v, Mapping) and isinstance(self._d.get(k), Conf):
                self._d[k].merge(v)
            else:
                self[k] = v
        return self

    def __repr__(self):
        return f"Conf({self.to_dict()})"


def _defaults_for_target(target: str) -> dict:
    """class `X` => dataclass `XConfig` in the same module (ganslate/configs/utils.py:55-61)."""
    module, name = target.rsplit(".", 1)
    try:
        cfg_cls = import_attr(f"{import_attr(target).__module__}.{name}Config")
    except (ImportError, AttributeError):
        return {}
    inst = cfg_cls(**{f.name: None for f in dataclasses.fields(cfg_cls)
                      if f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING})
    d = dataclasses.asdict(inst)
    return {k: v for k, v in d.items() if v != "???"}


def _complete(node: Any):
    if isinstance(node, dict):
        for k in list(node):
            node[k] = _complete(node[k])
        if isinstance(node.get("_target_"), str):
            merged = _defaults_for_target(node["_target_"])
            _deep_update(merged, node)
            return merged
    return node


def _deep_update(dst: dict, src: dict):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_update(dst[k], v)
        elif v is not None or k not in dst:
            dst[k] = v
    return dst


TRAIN_DEFAULTS = dict(
    output_dir="./outputs", batch_size=1, cuda=True, mixed_precision=False, opt_level="O1", n_iters=200000,
    n_iters_decay=0, seed=None, checkpointing=dict(load_iter=None, freq=2000, start_after=0, load_optimizers=True),
    logging=dict(freq=50), metrics=dict(discriminator_evolution=True, ssim=False))


def init_config(conf, overrides=()):
    """YAML path / dict -> Conf, with `_target_` nodes completed from their `XConfig` dataclasses and the
    reference's interpolation defaults (BA channels <- AB, discriminator A <- B; configs/base.py:30,42)."""
    if isinstance(conf, str):
        import yaml
        with open(conf) as f:
            conf = yaml.safe_load(f)
    conf = _complete(dict(conf))
    train = _deep_update({k: (dict(v) if isinstance(v, dict) else v) for k, v in TRAIN_DEFAULTS.items()},
                         conf.get("train", {}))
    conf["train"] = train
    conf.setdefault("mode", "train")
    gan = train.get("gan", {})
    ioc = gan.get("generator", {}).get("in_out_channels")
    if isinstance(ioc, dict) and ioc.get("BA") is None:
        ioc["BA"] = ioc.get("AB")
    ic = (gan.get("discriminator") or {}).get("in_channels")
    if isinstance(ic, dict) and ic.get("A") is None:
        ic["A"] = ic.get("B")
    c = Conf(conf)
    for item in overrides:  # dot-list overrides `a.b.c=value` (utils/builders.py:16-24)
        key, val = item.split("=", 1)
        import yaml
        node = c
        parts = key.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = yaml.safe_load(val)
    return c
