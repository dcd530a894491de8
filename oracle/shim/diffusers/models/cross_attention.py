class AttnProcessor:
    pass
