from parseq_b200.system import PARSeq, BatchResult  # noqa: F401
