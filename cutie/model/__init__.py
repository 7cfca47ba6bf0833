from cutie_b200.utils.shim_paths import compose_with_reference

__path__ = compose_with_reference(list(__path__), __name__)
