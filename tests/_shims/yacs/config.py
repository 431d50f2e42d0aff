"""Minimal yacs.config.CfgNode (attribute dict + merge_from_file/list + freeze) over PyYAML."""
import copy
import ast

import yaml


class CfgNode(dict):
    IMMUTABLE = "__immutable__"

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        init_dict = {} if init_dict is None else init_dict
        for k, v in list(init_dict.items()):
            if isinstance(v, dict) and not isinstance(v, CfgNode):
                init_dict[k] = CfgNode(v)
        super().__init__(init_dict)
        self.__dict__[CfgNode.IMMUTABLE] = False

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get(CfgNode.IMMUTABLE):
            raise AttributeError("Attempted to set {} to {}, but CfgNode is immutable".format(name, value))
        self[name] = value

    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def freeze(self):
        self._immutable(True)

    def defrost(self):
        self._immutable(False)

    def _immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(flag)

    def clone(self):
        return copy.deepcopy(self)

    def dump(self, **kw):
        def to_dict(n):
            return {k: to_dict(v) if isinstance(v, CfgNode) else v for k, v in n.items()}
        return yaml.safe_dump(to_dict(self), **kw)

    @staticmethod
    def _coerce(new, old, key):
        if isinstance(old, tuple) and isinstance(new, list):
            return tuple(new)
        if isinstance(old, list) and isinstance(new, tuple):
            return list(new)
        if isinstance(old, float) and isinstance(new, int):
            return float(new)
        if old is not None and new is not None and type(old) is not type(new) and not isinstance(old, CfgNode):
            raise ValueError("Type mismatch for {}: {} vs {}".format(key, type(old), type(new)))
        return new

    def _merge(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                raise KeyError("Non-existent config key: {}".format(full))
            if isinstance(v, dict):
                self[k]._merge(v, path + [k])
            else:
                if isinstance(v, str):
                    try:
                        v = ast.literal_eval(v)
                    except (ValueError, SyntaxError):
                        pass
                self[k] = CfgNode._coerce(v, self[k], full)

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {}, [])

    def merge_from_other_cfg(self, other):
        self._merge(other, [])

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for full, v in zip(lst[0::2], lst[1::2]):
            node = self
            keys = full.split(".")
            for k in keys[:-1]:
                node = node[k]
            if isinstance(v, str):
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            node[keys[-1]] = CfgNode._coerce(v, node[keys[-1]], full)

    def __deepcopy__(self, memo):
        n = CfgNode({k: copy.deepcopy(v, memo) for k, v in self.items()})
        n.__dict__[CfgNode.IMMUTABLE] = self.__dict__[CfgNode.IMMUTABLE]
        return n
