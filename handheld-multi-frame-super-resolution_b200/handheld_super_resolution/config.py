"""Configuration object.  The reference passes an OmegaConf DictConfig everywhere (run_handheld.py:94-116) and
`process()` mutates it; omegaconf is not installed in this image, so this module provides an attribute/item
mapping with the same access patterns the hot path uses (`cfg.a.b`, `cfg.a.get("k")`, `cfg.a.update({...})`).
A real OmegaConf object works too: the stage functions only use attribute access."""
import copy
import os

import yaml

DEFAULT_YAML = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "configs", "defaults.yaml")


class Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = Config.wrap(v)

    def __deepcopy__(self, memo):
        return Config({k: copy.deepcopy(v, memo) for k, v in self.items()})

    @staticmethod
    def wrap(d):
        if isinstance(d, Config):
            return d
        if isinstance(d, dict):
            return Config({k: Config.wrap(v) for k, v in d.items()})
        return d

    def merge_with(self, other):
        """Recursive update (OmegaConf.merge semantics for nested dicts)."""
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge_with(v)
            else:
                self[k] = Config.wrap(v)
        return self


def load_config(path=None, overrides=None):
    """configs/defaults.yaml merged with an optional user YAML and a dict of overrides (run_handheld.py:94-116)."""
    cfg = Config.wrap(yaml.safe_load(open(DEFAULT_YAML)))
    if path is not None:
        cfg.merge_with(yaml.safe_load(open(path)) or {})
    if overrides:
        cfg.merge_with(overrides)
    return cfg


def to_plain(cfg):
    """Nested plain dict (what the CPU oracle takes)."""
    if isinstance(cfg, dict):
        return {k: to_plain(v) for k, v in cfg.items()}
    if hasattr(cfg, "items") and not isinstance(cfg, (str, bytes)):
        return {k: to_plain(v) for k, v in cfg.items()}
    return cfg
