from cutie_b200.model.cutie import CUTIE  # noqa: F401
