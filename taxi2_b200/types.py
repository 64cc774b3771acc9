"""Boundary glue the reference takes from `itaxotools.common` (not vendored, not installable):
the `Type` subclass registry, plus `Container`, `Percentage` and `AttrDict`.

Behaviour follows /root/reference/src/itaxotools/taxi2/types.py:10-44 and the semantics pinned by
/root/reference/tests/test_types.py:8-37:
  * a subclass becomes an attribute of each of its DIRECT Type parents (PairwiseAligner.Biopython),
  * iterating a Type class yields its direct children, `Child in Parent` works on classes,
  * instances are not containers (`x in Parent()` raises TypeError) and compare equal by type.
"""
from __future__ import annotations

from typing import Callable, Generic, Iterable, Iterator, TypeVar

Item = TypeVar("Item")


class TypeMeta(type):
    """Metaclass keeping, per class, an ordered name -> class map of its direct subclasses."""

    def __init__(cls, name, bases, namespace, **kwargs):
        super().__init__(name, bases, namespace, **kwargs)
        cls._children = {}
        for base in bases:
            if isinstance(base, TypeMeta):
                base._children[name] = cls

    def __getattr__(cls, name):
        # only reached when normal lookup fails: resolve registered children by name
        children = cls.__dict__.get("_children", {})
        if name in children:
            return children[name]
        raise AttributeError(f"type object {cls.__name__!r} has no attribute {name!r}")

    def __dir__(cls):
        return list(super().__dir__()) + list(cls.__dict__.get("_children", {}))

    def __iter__(cls):
        return iter(cls.__dict__.get("_children", {}).values())

    def __contains__(cls, item) -> bool:
        return any(item is child for child in cls)

    def __len__(cls) -> int:
        return len(cls.__dict__.get("_children", {}))


class Type(metaclass=TypeMeta):
    """Base for registries such as PairwiseAligner, DistanceMetric, FileHandler."""

    def __eq__(self, other) -> bool:
        return type(self) is type(other)

    def __hash__(self) -> int:
        return hash(type(self))

    def __repr__(self) -> str:
        return f"<{type(self).__name__}>"

    @property
    def type(self):
        return type(self)


class Container(Generic[Item]):
    """Re-iterable wrapper around an iterable or a generator factory (types.py:10-39)."""

    def __init__(self, source: Iterable[Item] | Callable[..., Iterator[Item]], *args, **kwargs):
        if callable(source):
            self.callable, self.iterable = source, None
            self.args, self.kwargs = args, kwargs
        else:
            if args or kwargs:
                raise TypeError("Cannot pass arguments to iterable source")
            self.callable, self.iterable = None, source
            self.args, self.kwargs = (), {}

    def __iter__(self) -> Iterator[Item]:
        if self.callable is not None:
            return iter(self.callable(*self.args, **self.kwargs))
        return iter(self.iterable)

    def __len__(self) -> int:
        count = 0
        for _ in self:
            count += 1
        return count


class Percentage(float):
    def __str__(self) -> str:
        return f"{100 * self:.2f}%"


class AttrDict(dict):
    """dict with attribute access (itaxotools.common.utility.AttrDict, used for task params)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        try:
            del self[name]
        except KeyError as e:
            raise AttributeError(name) from e
