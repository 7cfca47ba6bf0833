from cutie_b200.inference.object_manager import ObjectInfo  # noqa: F401
