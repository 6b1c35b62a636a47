class VisionTransformer:  # imported but unused by the reference
    pass
