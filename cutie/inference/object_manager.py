from cutie_b200.inference.object_manager import ObjectManager  # noqa: F401
