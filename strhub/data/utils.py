from parseq_b200.tokenizer import CharsetAdapter, Tokenizer  # noqa: F401
