"""cattrs.gen stand-in (TEST INFRASTRUCTURE ONLY): hook generators used by baybe/serialization."""
from __future__ import annotations

import attrs


def override(*args, **kwargs):
    from cattrs import override as _o

    return _o(*args, **kwargs)


def make_dict_unstructure_fn(cl, converter, _cattrs_omit_if_default=False, **overrides):
    def unstructure(obj):
        out = {}
        for a in attrs.fields(type(obj)):
            ov = overrides.get(a.name)
            if ov is not None and getattr(ov, "omit", None):
                continue
            if not a.init and ov is None:
                continue
            name = (getattr(ov, "rename", None) if ov is not None else None) or a.name.lstrip("_")
            val = getattr(obj, a.name)
            hook = getattr(ov, "unstruct_hook", None) if ov is not None else None
            out[name] = hook(val) if hook else converter.unstructure(val)
        return out

    return unstructure


def make_dict_structure_fn(cl, converter, _cattrs_forbid_extra_keys=False, **overrides):
    def structure(obj, _tp=None):
        return converter.structure_attrs_fromdict(dict(obj), cl)

    return structure
