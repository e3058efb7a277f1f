"""`Sample`: the attribute-dict the visual adapters return (reference open_clip/util/Sample.py:21-53);
VisionTransformer.forward accepts a Tensor or any mapping with "x" and optional "pos" (transformer.py:732-745)."""
import collections.abc
from collections import OrderedDict


class Sample(OrderedDict):
    def __init__(self, init_dict=None):
        super().__init__(init_dict or {})

    def __setattr__(self, key, value):
        self[key] = value

    def __setitem__(self, key, value):
        if isinstance(value, collections.abc.Mapping) and not isinstance(value, Sample):
            value = Sample(value)
        super().__setitem__(key, value)

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def fields(self):
        return list(self.keys())
