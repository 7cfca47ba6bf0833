from cutie_b200.utils.tensor_utils import aggregate, pad_divide_by, unpad  # noqa: F401
