from cutie_b200.inference.memory_manager import MemoryManager  # noqa: F401
