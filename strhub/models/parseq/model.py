from parseq_b200.system import ParseqModel as PARSeq  # noqa: F401
