class DualTransformer2DModel:
    pass
