"""Minimal stand-in for the ``cattrs`` package -- TEST INFRASTRUCTURE ONLY.

The reference (``/root/reference/baybe``) imports cattrs at module import time for its (de)serialisation layer;
cattrs is not installed in this image and cannot be installed offline.  ``Campaign.recommend()`` itself never
(de)serialises, so this shim only has to (a) let ``import baybe`` succeed -- hook registration becomes
book-keeping -- and (b) structure the few plain containers BayBE converts at attribute-conversion time
(``cattrs.structure(x, tuple[float, ...])`` and friends).  Nothing under ``baybe_b200/`` imports it.
"""
from __future__ import annotations

import types
import typing
from typing import Any, get_args, get_origin

import attrs


class StructureHandlerNotFoundError(Exception):
    def __init__(self, message="", type_=None):
        super().__init__(message)
        self.type_ = type_


class BaseValidationError(ExceptionGroup if hasattr(__builtins__, "ExceptionGroup") or True else Exception):  # type: ignore[misc]
    def __new__(cls, message="", excs=None, cl=None):
        excs = list(excs) if excs else [ValueError(message or "validation error")]
        obj = super().__new__(cls, message, excs)
        obj.cl = cl
        return obj

    def derive(self, excs):
        return type(self)(self.message, excs, getattr(self, "cl", None))


class IterableValidationError(BaseValidationError):
    pass


class ClassValidationError(BaseValidationError):
    pass


class ForbiddenExtraKeysError(Exception):
    pass


def _is_union(tp) -> bool:
    return get_origin(tp) in (typing.Union, types.UnionType)


def _structure_plain(obj: Any, tp: Any, conv: "Converter | None" = None) -> Any:
    """Structure basic containers / scalars / attrs classes from already-python data."""
    if tp is Any or tp is None or tp is type(None):
        return obj
    if conv is not None:
        hook = conv._find_structure_hook(tp)
        if hook is not None:
            return hook(obj, tp)
    origin = get_origin(tp)
    if _is_union(tp):
        args = get_args(tp)
        if obj is None and type(None) in args:
            return None
        for a in args:
            if a is type(None):
                continue
            try:
                if isinstance(a, type) and isinstance(obj, a):
                    return obj
            except TypeError:
                pass
        last = None
        for a in args:
            if a is type(None):
                continue
            try:
                return _structure_plain(obj, a, conv)
            except Exception as e:  # noqa: BLE001
                last = e
        raise last if last else StructureHandlerNotFoundError(f"cannot structure {obj!r} as {tp}", tp)
    if origin in (tuple,):
        args = get_args(tp)
        if len(args) == 2 and args[1] is Ellipsis:
            return tuple(_structure_plain(v, args[0], conv) for v in obj)
        if args:
            return tuple(_structure_plain(v, a, conv) for v, a in zip(obj, args))
        return tuple(obj)
    if origin in (list, typing.List) or tp is list:
        (a,) = get_args(tp) or (Any,)
        return [_structure_plain(v, a, conv) for v in obj]
    if origin in (set, frozenset):
        (a,) = get_args(tp) or (Any,)
        return origin(_structure_plain(v, a, conv) for v in obj)
    if origin in (dict,):
        ka, va = get_args(tp) or (Any, Any)
        return {_structure_plain(k, ka, conv): _structure_plain(v, va, conv) for k, v in obj.items()}
    if origin is typing.Literal:
        if obj in get_args(tp):
            return obj
        raise ValueError(f"{obj!r} not in {get_args(tp)}")
    if tp in (float, int, str, bool, bytes):
        if tp is float:
            return float(obj)
        if tp is int:
            return int(obj)
        if tp is bool:
            return bool(obj)
        return tp(obj)
    if isinstance(tp, type):
        if isinstance(obj, tp):
            return obj
        if attrs.has(tp) and isinstance(obj, dict):
            return conv.structure_attrs_fromdict(obj, tp) if conv else tp(**obj)
        import enum

        if issubclass(tp, enum.Enum):
            return tp(obj)
    raise StructureHandlerNotFoundError(f"Unsupported type: {tp!r} (cattrs test shim)", tp)


def _unstructure_plain(obj: Any, conv: "Converter | None" = None) -> Any:
    import enum

    if conv is not None:
        hook = conv._find_unstructure_hook(type(obj))
        if hook is not None:
            return hook(obj)
    if isinstance(obj, enum.Enum):
        return obj.value
    if attrs.has(type(obj)):
        return {a.name.lstrip("_"): _unstructure_plain(getattr(obj, a.name), conv)
                for a in attrs.fields(type(obj)) if a.init}
    if isinstance(obj, dict):
        return {_unstructure_plain(k, conv): _unstructure_plain(v, conv) for k, v in obj.items()}
    if isinstance(obj, (list, tuple, set, frozenset)):
        return [_unstructure_plain(v, conv) for v in obj]
    return obj


class Converter:
    """Book-keeping converter: registration calls are recorded, exact-type and predicate hooks are honoured."""

    def __init__(self, *args, **kwargs):
        self._s_hooks: dict = {}
        self._s_funcs: list = []
        self._s_factories: list = []
        self._u_hooks: dict = {}
        self._u_funcs: list = []
        self._u_factories: list = []
        self.forbid_extra_keys = kwargs.get("forbid_extra_keys", False)

    # -- registration -------------------------------------------------------------------------------
    def register_structure_hook(self, cl=None, func=None):
        if func is None and cl is not None and callable(cl) and not isinstance(cl, type) and get_origin(cl) is None:
            f = cl  # decorator form: the class is the annotation of the return value
            hints = typing.get_type_hints(f)
            tp = hints.get("return")
            self._s_hooks[tp] = f
            return f
        if func is None:
            def deco(f):
                self._s_hooks[cl] = f
                return f
            return deco
        self._s_hooks[cl] = func
        return func

    def register_unstructure_hook(self, cl=None, func=None):
        if func is None and cl is not None and callable(cl) and not isinstance(cl, type) and get_origin(cl) is None:
            f = cl
            hints = typing.get_type_hints(f)
            hints.pop("return", None)
            tp = next(iter(hints.values()), None)
            self._u_hooks[tp] = f
            return f
        if func is None:
            def deco(f):
                self._u_hooks[cl] = f
                return f
            return deco
        self._u_hooks[cl] = func
        return func

    def register_structure_hook_func(self, check, func):
        self._s_funcs.insert(0, (check, func))

    def register_unstructure_hook_func(self, check, func):
        self._u_funcs.insert(0, (check, func))

    def register_structure_hook_factory(self, check, factory=None):
        if factory is None:
            def deco(f):
                self._s_factories.insert(0, (check, f))
                return f
            return deco
        self._s_factories.insert(0, (check, factory))
        return factory

    def register_unstructure_hook_factory(self, check, factory=None):
        if factory is None:
            def deco(f):
                self._u_factories.insert(0, (check, f))
                return f
            return deco
        self._u_factories.insert(0, (check, factory))
        return factory

    # -- lookup -------------------------------------------------------------------------------------
    def _find_structure_hook(self, tp):
        try:
            if tp in self._s_hooks:
                return self._s_hooks[tp]
        except TypeError:
            pass
        for check, func in self._s_funcs:
            try:
                if check(tp):
                    return func
            except Exception:  # noqa: BLE001
                continue
        for check, fac in self._s_factories:
            try:
                if check(tp):
                    try:
                        return fac(tp)
                    except TypeError:
                        return fac(tp, self)
            except Exception:  # noqa: BLE001
                continue
        return None

    def _find_unstructure_hook(self, tp):
        try:
            if tp in self._u_hooks:
                return self._u_hooks[tp]
        except TypeError:
            pass
        for check, func in self._u_funcs:
            try:
                if check(tp):
                    return func
            except Exception:  # noqa: BLE001
                continue
        for check, fac in self._u_factories:
            try:
                if check(tp):
                    try:
                        return fac(tp)
                    except TypeError:
                        return fac(tp, self)
            except Exception:  # noqa: BLE001
                continue
        return None

    def get_structure_hook(self, tp, cache_result=True):
        hook = self._find_structure_hook(tp)
        return hook if hook is not None else (lambda obj, t=tp: _structure_plain(obj, t, None))

    def get_unstructure_hook(self, tp, cache_result=True):
        hook = self._find_unstructure_hook(tp)
        return hook if hook is not None else (lambda obj: _unstructure_plain(obj, None))

    # -- conversion ---------------------------------------------------------------------------------
    def structure(self, obj, cl):
        return _structure_plain(obj, cl, self)

    def unstructure(self, obj, unstructure_as=None):
        return _unstructure_plain(obj, self)

    def structure_attrs_fromdict(self, obj, cl):
        hints = typing.get_type_hints(cl)
        kwargs = {}
        for a in attrs.fields(cl):
            if not a.init:
                continue
            name = a.alias or a.name.lstrip("_")
            if name in obj:
                kwargs[name] = _structure_plain(obj[name], hints.get(a.name, Any), self)
        return cl(**kwargs)

    def unstructure_attrs_asdict(self, obj):
        return _unstructure_plain(obj, None)

    def copy(self, **kwargs):
        new = Converter(forbid_extra_keys=kwargs.get("forbid_extra_keys", self.forbid_extra_keys))
        new._s_hooks = dict(self._s_hooks)
        new._s_funcs = list(self._s_funcs)
        new._s_factories = list(self._s_factories)
        new._u_hooks = dict(self._u_hooks)
        new._u_funcs = list(self._u_funcs)
        new._u_factories = list(self._u_factories)
        return new


BaseConverter = Converter
global_converter = Converter()


def structure(obj, cl):
    return global_converter.structure(obj, cl)


def unstructure(obj, unstructure_as=None):
    return global_converter.unstructure(obj, unstructure_as)


def register_structure_hook(cl, func=None):
    return global_converter.register_structure_hook(cl, func)


def register_unstructure_hook(cl, func=None):
    return global_converter.register_unstructure_hook(cl, func)


class _Override:
    def __init__(self, omit_if_default=None, rename=None, omit=None, struct_hook=None, unstruct_hook=None):
        self.omit_if_default, self.rename, self.omit = omit_if_default, rename, omit
        self.struct_hook, self.unstruct_hook = struct_hook, unstruct_hook


def override(omit_if_default=None, rename=None, omit=None, struct_hook=None, unstruct_hook=None):
    return _Override(omit_if_default, rename, omit, struct_hook, unstruct_hook)


from cattrs import dispatch, errors, gen, strategies  # noqa: E402,F401
