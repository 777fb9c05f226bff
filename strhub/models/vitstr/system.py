from parseq_b200.system import ViTSTR, BatchResult  # noqa: F401
