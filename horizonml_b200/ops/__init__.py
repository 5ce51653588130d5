"""Fused op set: ``functional`` (autograd layer), ``torch_backend`` (oracle / CPU path),
``native_backend`` (hand-written sm_100a kernels in ``csrc/``)."""
from .functional import (  # noqa: F401
    set_backend, get_backend, conv_bn_act, dwconv_bn_act, maxpool3x3s2, head_loss, head_logits,
    adam_step, grad_diff_sq, stem_prepare, compute_weight, grad_target, grad_written,
    enable_side_stream, join_side, step_begin, step_end, stats_update, GradLink, BNBackLink, backward,
)


def native_available() -> bool:
    """True when the CUDA extension is built and a GPU is visible."""
    import torch
    if not torch.cuda.is_available():
        return False
    try:
        from . import _ext
        return _ext.load(required=False) is not None
    except Exception:
        return False
