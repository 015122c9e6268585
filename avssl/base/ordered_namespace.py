"""Config container with dict + attribute access (reference: avssl/base/ordered_namespace.py:7-153).

The import path ``avssl.base.ordered_namespace.OrderedNamespace`` is part of the reference's checkpoint format:
``save_hyperparameters()`` pickles the config into every ``.ckpt`` (avssl/model/base_model.py:15), and the pickle
stores the instance ``__dict__`` (an OrderedDict of entries), restored through ``__setstate__``.
"""
from argparse import Namespace
from collections import OrderedDict
from types import SimpleNamespace

_NS = (SimpleNamespace, Namespace)


def _wrap(value):
    if isinstance(value, dict):
        return OrderedNamespace(value)
    if isinstance(value, _NS):
        return OrderedNamespace(vars(value))
    if isinstance(value, list):
        return [OrderedNamespace(v) if isinstance(v, dict) else v for v in value]
    return value


class OrderedNamespace(object):
    def __init__(self, data=None, **kwargs):
        object.__setattr__(self, "_odict", OrderedDict())
        if data is None:
            sources = [kwargs]
        elif isinstance(data, (tuple, list)):
            sources = list(data)  # merged left to right: later entries win
        else:
            sources = [data]
        for src in sources:
            if isinstance(src, OrderedNamespace):
                src = src._odict
            elif isinstance(src, _NS):
                src = vars(src)
            for key, value in src.items():
                self._odict[key] = _wrap(value)

    # attribute / item access share one store
    def __getattr__(self, key):
        store = object.__getattribute__(self, "_odict")
        try:
            return store[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self._odict[key] = value

    def __getitem__(self, key):
        return self.__getattr__(key)

    def __setitem__(self, key, value):
        self._odict[key] = value

    def __delitem__(self, key):
        del self._odict[key]

    def __contains__(self, key):
        return key in self._odict

    def __iter__(self):
        return iter(self.to_dict())

    def __len__(self):
        return len(self._odict)

    def __eq__(self, other):
        return isinstance(other, OrderedNamespace) and self._odict == other._odict

    def __ne__(self, other):
        return not self.__eq__(other)

    @property
    def __dict__(self):
        return self._odict

    def __getstate__(self):
        return self._odict

    def __setstate__(self, state):
        object.__setattr__(self, "_odict", OrderedDict())
        self._odict.update(state)

    def _convert(self, factory):
        out = factory()
        for key, value in self._odict.items():
            out[key] = value._convert(factory) if isinstance(value, OrderedNamespace) else value
        return out

    def to_odict(self):
        return self._convert(OrderedDict)

    def to_dict(self):
        return self._convert(dict)

    odict = property(to_odict)
    pydict = property(to_dict)

    def keys(self):
        return self._odict.keys()

    def items(self):
        return self._odict.items()

    def values(self):
        return self._odict.values()

    def get(self, key, value=None):
        return self._odict.get(key, value)

    def copy(self):
        return self.__class__(self)

    def __str__(self):
        return "OrderedNamespace(" + str(self.to_dict()) + ")"

    __repr__ = __str__
