from clipself_b200.training.clipself import CLIPSelf  # noqa: F401
