"""Minimal stand-in for the `optree` package (absent from this image, no network).

TEST INFRASTRUCTURE ONLY: lets tests/golden/make_golden.py import the read-only
reference at /root/reference so golden vectors can be generated in the build
container.  Never imported by cola_b200/ or by anything that runs on the GPU box.
Covers only the calls the reference makes (cola/backends/torch_fns.py:329-338,
cola/backends/backends.py:81-83).
"""
_REGISTRY = {}


def register_pytree_node_class(cls=None, namespace=None):
    def deco(c):
        _REGISTRY[c] = namespace
        return c
    return deco if cls is None else deco(cls)


class _Spec:
    def __init__(self, kind, meta, children):
        self.kind, self.meta, self.children = kind, meta, children

    @property
    def num_leaves(self):
        return 1 if self.kind == "leaf" else sum(c.num_leaves for c in self.children)


def _registered(obj):
    for klass in type(obj).__mro__:
        if klass in _REGISTRY:
            return True
    return False


def _flatten(obj, leaves):
    if obj is None:
        return _Spec("none", None, [])
    if isinstance(obj, (list, tuple)) and not hasattr(obj, "_fields"):
        return _Spec("seq", type(obj), [_flatten(o, leaves) for o in obj])
    if isinstance(obj, dict):
        keys = sorted(obj.keys(), key=str)
        return _Spec("dict", keys, [_flatten(obj[k], leaves) for k in keys])
    if _registered(obj):
        children, aux = obj.tree_flatten()
        return _Spec("node", (type(obj), aux), [_flatten(c, leaves) for c in children])
    leaves.append(obj)
    return _Spec("leaf", None, [])


def tree_flatten(tree, namespace=None, **_):
    leaves = []
    spec = _flatten(tree, leaves)
    return leaves, spec


def tree_structure(tree, namespace=None, **_):
    return tree_flatten(tree)[1]


def treespec_is_leaf(spec):
    return spec.kind == "leaf"


def _unflatten(spec, it):
    if spec.kind == "leaf":
        return next(it)
    if spec.kind == "none":
        return None
    kids = [_unflatten(c, it) for c in spec.children]
    if spec.kind == "seq":
        return spec.meta(kids)
    if spec.kind == "dict":
        return dict(zip(spec.meta, kids))
    klass, aux = spec.meta
    return klass.tree_unflatten(aux, kids)


def tree_unflatten(spec, leaves):
    return _unflatten(spec, iter(leaves))
