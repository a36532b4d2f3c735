"""Minimal ``opt.yaml`` config object (OmegaConf is what the reference uses, ``denoise_object.py:33-37``; it is not
a dependency here).  Supports what the hot path needs from a config node: attribute access, ``in``, ``.get``,
item assignment, dotted CLI overrides and merging."""
from __future__ import annotations

import copy
from typing import Any

import yaml


class Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = Config.wrap(v)

    def __deepcopy__(self, memo):
        return Config.wrap(copy.deepcopy(dict(self), memo))

    @staticmethod
    def wrap(d: Any):
        if isinstance(d, Config):
            return d
        if isinstance(d, dict):
            return Config({k: Config.wrap(v) for k, v in d.items()})
        if isinstance(d, (list, tuple)):
            return [Config.wrap(v) for v in d]
        return d

    def to_dict(self) -> dict:
        def un(v):
            if isinstance(v, dict):
                return {k: un(x) for k, x in v.items()}
            if isinstance(v, list):
                return [un(x) for x in v]
            return v

        return un(self)

    def merge(self, other: dict) -> "Config":
        """Recursive merge (``OmegaConf.merge`` semantics for dicts: right side wins)."""
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge(v)
            else:
                self[k] = Config.wrap(v)
        return self

    def set_dotted(self, key: str, value: Any) -> None:
        node = self
        ks = key.split(".")
        for k in ks[:-1]:
            if k not in node or not isinstance(node[k], dict):
                node[k] = Config()
            node = node[k]
        node[ks[-1]] = Config.wrap(value)


def load_yaml(path: str) -> Config:
    with open(path) as f:
        return Config.wrap(yaml.safe_load(f))


def to_plain(cfg) -> dict:
    return cfg.to_dict() if isinstance(cfg, Config) else Config.wrap(cfg).to_dict()
