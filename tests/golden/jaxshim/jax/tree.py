"""Stand-in for jax.tree over tuples / lists / dicts."""


def map(f, tree, *rest):
    if isinstance(tree, tuple):
        return tuple(map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
    if isinstance(tree, list):
        return [map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree)]
    if isinstance(tree, dict):
        return {k: map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if tree is None:
        return None
    return f(tree, *rest)


def leaves(tree):
    if isinstance(tree, (tuple, list)):
        return [l for t in tree for l in leaves(t)]
    if isinstance(tree, dict):
        return [l for t in tree.values() for l in leaves(t)]
    if tree is None:
        return []
    return [tree]
