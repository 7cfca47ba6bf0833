from cutie_b200.inference.inference_core import InferenceCore  # noqa: F401
