from cutie_b200.inference.memory_bank import KeyValueMemoryStore  # noqa: F401
