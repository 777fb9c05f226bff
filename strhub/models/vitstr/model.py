from parseq_b200.system import VitstrModel as ViTSTR  # noqa: F401
