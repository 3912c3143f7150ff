"""A minimal stand-in for the ``parameterized`` package (absent from this image) covering what the reference's unit
tests use: ``@parameterized.expand(cases, name_func=...)`` on TestCase methods and ``@parameterized_class``."""
import functools
import inspect


class param:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


def _as_param(case):
    if isinstance(case, param):
        return case
    if isinstance(case, (tuple, list)):
        return param(*case)
    return param(case)


class parameterized:
    @classmethod
    def expand(cls, cases, name_func=None, **_ignored):
        cases = [_as_param(c) for c in cases]

        def decorate(fn):
            namespace = inspect.currentframe().f_back.f_locals  # the class body being executed
            for i, p in enumerate(cases):
                name = name_func(fn, i, p) if name_func else f"{fn.__name__}_{i}"
                while name in namespace:
                    name += f"_{i}"

                def bind(p):
                    @functools.wraps(fn)
                    def test(self):
                        return fn(self, *p.args, **p.kwargs)

                    return test

                test = bind(p)
                test.__name__ = name
                namespace[name] = test
            return None  # the un-expanded method disappears from the class

        return decorate


def parameterized_class(attrs, values=None, class_name_func=None, **_ignored):
    if values is None:  # list-of-dicts form
        dicts = list(attrs)
    else:
        names = (attrs,) if isinstance(attrs, str) else tuple(attrs)
        dicts = [dict(zip(names, v)) for v in values]

    def decorate(base):
        module = inspect.getmodule(base).__dict__
        tests = {k: v for k, v in vars(base).items() if k.startswith("test")}
        for i, d in enumerate(dicts):
            name = class_name_func(base, i, d) if class_name_func else f"{base.__name__}_{i}"
            module[name] = type(name, (base,), {**tests, **d})
        for key in tests:
            delattr(base, key)  # only the generated subclasses run
        return base

    return decorate
