"""Shared by the `cutie.*` shim packages: lets sub-modules this shim does not provide resolve to a reference checkout
placed behind this repo on sys.path."""
import os
import sys


def compose_with_reference(path, name):
    """If the reference checkout is also on sys.path (behind this repo), let sub-modules this shim does not provide
    -- dataset readers, result savers, palette ... -- resolve to the reference's files, so that `eval_vos.py`-style
    callers import unchanged; modules provided here (the hot-path surface) keep precedence."""
    sub = name.replace('.', os.sep)
    have = {os.path.abspath(p) for p in path}
    for root in sys.path:
        d = os.path.abspath(os.path.join(root or '.', sub))
        if os.path.isdir(d) and d not in have:
            path.append(d)
            have.add(d)
    return path
