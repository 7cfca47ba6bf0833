from cutie_b200.inference.image_feature_store import ImageFeatureStore  # noqa: F401
