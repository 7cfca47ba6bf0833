"""Drop-in import surface: `cutie.*` names of the reference (hkchengrex/Cutie) resolved to the B200-native
implementation in cutie_b200, so scripting_demo.py / scripting_demo_add_del_objects.py / process_video.py
style callers (`from cutie.inference.inference_core import InferenceCore`) run unchanged.
Only the hot-path surface is provided; dataset readers, result savers, training
and GUI modules of the reference are out of scope."""


def _compose_with_reference(path, name):
    """If the reference checkout is also on sys.path (behind this repo), let sub-modules this shim does not provide
    -- dataset readers, result savers, palette ... -- resolve to the reference's files, so that `eval_vos.py`-style
    callers import unchanged; modules provided here (the hot-path surface) keep precedence."""
    import os
    import sys
    sub = name.replace('.', os.sep)
    have = {os.path.abspath(p) for p in path}
    for root in sys.path:
        d = os.path.abspath(os.path.join(root or '.', sub))
        if os.path.isdir(d) and d not in have:
            path.append(d)
            have.add(d)
    return path


__path__ = _compose_with_reference(list(__path__), __name__)
