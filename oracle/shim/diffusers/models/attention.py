class AdaGroupNorm:
    pass


class AttentionBlock:
    pass
