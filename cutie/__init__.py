"""Drop-in import surface: `cutie.*` names of the reference (hkchengrex/Cutie) resolved to the B200-native
implementation in cutie_b200, so scripting_demo.py / scripting_demo_add_del_objects.py / process_video.py
style callers (`from cutie.inference.inference_core import InferenceCore`) run unchanged.
Only the hot-path surface is provided; dataset readers, result savers, training
and GUI modules of the reference are out of scope."""
from cutie_b200.utils.shim_paths import compose_with_reference

__path__ = compose_with_reference(list(__path__), __name__)
