from parseq_b200.factory import create_model, load_from_checkpoint, parse_model_args, get_pretrained_weights  # noqa: F401
from parseq_b200.system import InvalidModelError  # noqa: F401
