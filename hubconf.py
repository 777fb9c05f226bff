"""torch.hub entry points with the reference's names and signatures (hubconf.py:6-33) for the PARSeq
family and ViTSTR (hubconf.py:53-58).  Unlike the reference, no pytorch_lightning / timm dependency."""
from parseq_b200.factory import create_model

dependencies = ['torch']


def parseq_tiny(pretrained: bool = False, decode_ar: bool = True, refine_iters: int = 1, **kwargs):
    """PARSeq-Ti: 32x128 crops, 4x8 patches, d_model=192."""
    return create_model('parseq-tiny', pretrained, decode_ar=decode_ar, refine_iters=refine_iters, **kwargs)


def parseq(pretrained: bool = False, decode_ar: bool = True, refine_iters: int = 1, **kwargs):
    """PARSeq-S: 32x128 crops, 4x8 patches, d_model=384."""
    return create_model('parseq', pretrained, decode_ar=decode_ar, refine_iters=refine_iters, **kwargs)


def parseq_patch16_224(pretrained: bool = False, decode_ar: bool = True, refine_iters: int = 1, **kwargs):
    """PARSeq-S on 224x224 crops with 16x16 patches."""
    return create_model('parseq-patch16-224', pretrained, decode_ar=decode_ar, refine_iters=refine_iters, **kwargs)


def vitstr(pretrained: bool = False, **kwargs):
    """ViTSTR-S: 32x128 crops, 4x8 patches, d_model=384."""
    return create_model('vitstr', pretrained, **kwargs)
