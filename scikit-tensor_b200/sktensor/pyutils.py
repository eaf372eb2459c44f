"""Small host-side helpers with the reference's names (sktensor/pyutils.py)."""
import numbers
from collections.abc import Sequence


def inherit_docstring_from(cls):
    """Decorator: take the docstring of the same-named attribute of ``cls`` (pyutils.py:1-5)."""
    def deco(fn):
        fn.__doc__ = getattr(getattr(cls, fn.__name__, None), '__doc__', None)
        return fn
    return deco


def is_sequence(obj):
    """True for list/tuple-like containers (pyutils.py:8-19)."""
    return isinstance(obj, Sequence)


def is_number(obj):
    """True for Python / NumPy scalars (pyutils.py:22-33)."""
    return isinstance(obj, numbers.Number)


def func_attr(f, attr):
    """``f.__<attr>__`` with the Python-2 spelling as a fallback (pyutils.py:36-46)."""
    for name in ('__%s__' % attr, 'func_%s' % attr):
        if hasattr(f, name):
            return getattr(f, name)
    raise ValueError('Object %s has no attr %s' % (str(f), attr))


def from_to_without(frm, to, without, step=1, skip=1, reverse=False, separate=False):
    """``range(frm, to, step)`` with the entry ``without`` left out (pyutils.py:49-62).

    ``reverse=True`` walks ``to-1 .. frm`` instead; ``separate=True`` returns the part
    before and the part after the gap as two lists.
    """
    if reverse:
        frm, to = to - 1, frm - 1
        step, skip = -step, -skip
    head = list(range(frm, without, step))
    tail = list(range(without + skip, to, step))
    return (head, tail) if separate else head + tail
