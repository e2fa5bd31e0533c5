"""Minimal yacs-compatible CfgNode: nested attribute access, merge_from_file (YAML), merge_from_list (literal-eval'd KEY VALUE
pairs), freeze / defrost, dump, clone, str - what os2d/config.py, main.py:33-35 and os2d/utils/logger.py:112 call."""
import ast
import copy

import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get("_frozen", False):
            raise AttributeError("Attempted to set {} to {}, but CfgNode is immutable".format(name, value))
        self[name] = value

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def _set_frozen(self, flag):
        self.__dict__["_frozen"] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def is_frozen(self):
        return self.__dict__["_frozen"]

    def clone(self):
        return copy.deepcopy(self)

    def _to_dict(self):
        return {k: (v._to_dict() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def dump(self, **kwargs):
        return yaml.safe_dump(self._to_dict(), **kwargs)

    def __str__(self):
        return self.dump()

    __repr__ = __str__

    def _merge(self, other):
        for k, v in other.items():
            if k not in self:
                raise KeyError("Non-existent config key: {}".format(k))
            if isinstance(self[k], CfgNode) and isinstance(v, dict):
                self[k]._merge(v)
            else:
                self[k] = v

    def merge_from_file(self, path):
        with open(path, "r") as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_other_cfg(self, other):
        self._merge(other._to_dict())

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0, "Override list has odd length: {}".format(cfg_list)
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            keys = full_key.split(".")
            for k in keys[:-1]:
                node = node[k]
            if keys[-1] not in node:
                raise KeyError("Non-existent config key: {}".format(full_key))
            if isinstance(v, str):
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            node[keys[-1]] = v
