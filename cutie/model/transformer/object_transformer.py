from cutie_b200.model.object_transformer import QueryTransformer, QueryTransformerBlock  # noqa: F401
