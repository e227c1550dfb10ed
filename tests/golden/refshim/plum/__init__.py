"""Minimal stand-in for `cola-plum-dispatch` (absent from this image, no network).

TEST INFRASTRUCTURE ONLY: lets tests/golden/make_golden.py import the read-only
reference at /root/reference so golden vectors can be generated in the build
container.  Never imported by cola_b200/ or by anything that runs on the GPU box.

Implements the subset of plum the reference uses: `@dispatch`,
`@dispatch(precedence=, cond=)`, `@dispatch.abstract` (defaults of the abstract
signature are applied at call time, reference cola/linalg/inverse/inv.py:42-44),
and `@parametric` (covariant per-argument-type subclasses such as
`Product[Dense, Dense]`, reference cola/annotations.py:97-98,116).
"""
import inspect
import types
import typing

_REG = {}


def _is_union(hint):
    origin = typing.get_origin(hint)
    return origin is typing.Union or origin is types.UnionType


def _hint_options(hint):
    if _is_union(hint):
        out = []
        for h in typing.get_args(hint):
            out.extend(_hint_options(h))
        return out
    return [hint]


def _isinstance(obj, hint):
    if hint is inspect.Parameter.empty or hint is typing.Any:
        return True
    for h in _hint_options(hint):
        if h is typing.Any:
            return True
        if h is None or h is type(None):
            if obj is None:
                return True
            continue
        origin = typing.get_origin(h)
        if origin is not None:  # List[...], Callable[...], Set[...]: check the origin only
            import collections.abc as cabc
            if origin is cabc.Callable:
                if callable(obj):
                    return True
                continue
            h = origin
        if h is typing.Callable:
            if callable(obj):
                return True
            continue
        try:
            if isinstance(obj, h):
                return True
        except TypeError:
            pass
    return False


def _issubhint(a, b):
    """True when every value matching hint `a` also matches hint `b`."""
    if b is inspect.Parameter.empty or b is typing.Any:
        return True
    if a is inspect.Parameter.empty or a is typing.Any:
        return False
    for ha in _hint_options(a):
        ok = False
        for hb in _hint_options(b):
            if hb is typing.Any:
                ok = True
                break
            oa, ob = typing.get_origin(ha) or ha, typing.get_origin(hb) or hb
            try:
                if oa is ob or issubclass(oa, ob):
                    ok = True
                    break
            except TypeError:
                if oa == ob:
                    ok = True
                    break
        if not ok:
            return False
    return True


class _Method:
    def __init__(self, f, precedence, cond):
        self.f, self.precedence, self.cond = f, precedence, cond
        self.sig = inspect.signature(f)
        try:
            hints = typing.get_type_hints(f)
        except Exception:
            hints = {}
        self.params = []
        self.has_var = False
        for name, p in self.sig.parameters.items():
            if p.kind in (p.VAR_POSITIONAL, ):
                self.has_var = True
                self.var_hint = hints.get(name, p.annotation)
                continue
            if p.kind in (p.VAR_KEYWORD, p.KEYWORD_ONLY):
                continue
            self.params.append((name, hints.get(name, p.annotation), p.default))

    def matches(self, args):
        if len(args) > len(self.params) and not self.has_var:
            return False
        for i, a in enumerate(args):
            if i < len(self.params):
                if not _isinstance(a, self.params[i][1]):
                    return False
            elif not _isinstance(a, self.var_hint):
                return False
        for name, hint, default in self.params[len(args):]:
            if default is inspect.Parameter.empty:
                return False
        return True

    def hints_for(self, nargs):
        return [self.params[i][1] if i < len(self.params) else self.var_hint for i in range(nargs)]


class _Function:
    def __init__(self, name):
        self.__name__ = name
        self.__qualname__ = name
        self.methods = []
        self.abstract_sig = None
        self.__doc__ = None
        self.__module__ = None

    def register(self, f, precedence=0, cond=None):
        self.methods.append(_Method(f, precedence, cond))
        if self.__doc__ is None:
            self.__doc__ = f.__doc__
        self.__module__ = f.__module__
        self.__wrapped_name__ = f.__name__
        return self

    def _normalise(self, args, kwargs):
        if self.abstract_sig is not None:
            try:
                bound = self.abstract_sig.bind(*args, **kwargs)
                bound.apply_defaults()
                return tuple(bound.args), dict(bound.kwargs)
            except TypeError:
                pass
        return args, kwargs

    def resolve(self, args):
        cands = [m for m in self.methods if m.matches(args)]
        cands = [m for m in cands if m.cond is None or m.cond(*args)]
        if not cands:
            raise LookupError(f"no method of {self.__name__} for {tuple(type(a).__name__ for a in args)}")
        n = len(args)

        def dominates(m1, m2):  # m1 at least as specific as m2 everywhere, and differs
            h1, h2 = m1.hints_for(n), m2.hints_for(n)
            le = all(_issubhint(a, b) for a, b in zip(h1, h2))
            ge = all(_issubhint(b, a) for a, b in zip(h1, h2))
            return le and not ge

        best = [m for m in cands if not any(dominates(o, m) for o in cands if o is not m)]
        best.sort(key=lambda m: (m.cond is not None, m.precedence), reverse=True)
        return best[0]

    def __call__(self, *args, **kwargs):
        args, kwargs = self._normalise(args, kwargs)
        pos = list(args)
        # keyword arguments that name positional parameters take part in dispatch
        if kwargs and self.methods:
            names = [p[0] for p in self.methods[0].params]
            for name in names[len(pos):]:
                if name in kwargs:
                    pos.append(kwargs.pop(name))
                else:
                    break
        method = self.resolve(tuple(pos))
        return method.f(*pos, **kwargs)

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        return types.MethodType(self, obj)


class _Dispatcher:
    def _get(self, f):
        fn = _REG.get(f.__name__)
        if fn is None:
            fn = _REG[f.__name__] = _Function(f.__name__)
        fn.__module__ = f.__module__
        return fn

    def __call__(self, f=None, precedence=0, cond=None):
        if f is None:
            return lambda g: self._get(g).register(g, precedence=precedence, cond=cond)
        return self._get(f).register(f)

    def abstract(self, f):
        fn = self._get(f)
        fn.abstract_sig = inspect.signature(f)
        fn.__doc__ = f.__doc__
        return fn


dispatch = _Dispatcher()


def _param_sub(p, q):
    return _issubhint(p, q)


def parametric(cls):
    base_meta = type(cls)

    class ParametricMeta(base_meta):
        def __getitem__(c, params):
            root = c.__dict__.get("_p_root", None) or c
            if not isinstance(params, tuple):
                params = (params, )
            cache = root.__dict__["_p_cache"]
            key = tuple(params)
            try:
                hit = cache.get(key)
            except TypeError:
                hit, key = None, tuple(map(repr, params))
                hit = cache.get(key)
            if hit is None:
                name = f"{root.__name__}[{', '.join(getattr(p, '__name__', str(p)) for p in params)}]"
                hit = ParametricMeta(name, (root, ), {"_p_params": tuple(params), "_p_root": root,
                                                      "__module__": root.__module__})
                cache[key] = hit
            return hit

        def __call__(c, *args, **kwargs):
            if "_p_params" not in c.__dict__ and args:
                c = c[tuple(type(a) for a in args)]
            return super(ParametricMeta, c).__call__(*args, **kwargs)

        def __subclasscheck__(c, sub):
            cp = c.__dict__.get("_p_params")
            if cp is None:
                return base_meta.__subclasscheck__(c, sub)
            root = c.__dict__["_p_root"]
            if not isinstance(sub, type):
                return False
            sp = sub.__dict__.get("_p_params") if hasattr(sub, "__dict__") else None
            if sp is None:
                # subclasses of a concrete parametric type inherit its parameters
                for klass in getattr(sub, "__mro__", ()):
                    if "_p_params" in klass.__dict__ and klass.__dict__.get("_p_root") is root:
                        sp = klass.__dict__["_p_params"]
                        break
            if sp is None or not base_meta.__subclasscheck__(root, sub):
                return False
            if len(sp) != len(cp):
                return False
            return all(_param_sub(a, b) for a, b in zip(sp, cp))

        def __instancecheck__(c, inst):
            return c.__subclasscheck__(type(inst))

    new = ParametricMeta(cls.__name__, (cls, ), {"_p_cache": {}, "__module__": cls.__module__,
                                                 "__doc__": cls.__doc__, "__qualname__": cls.__qualname__})
    return new


__all__ = ["dispatch", "parametric"]
