from .vision_transformer import VisionTransformer, create_model, list_models  # noqa: F401
