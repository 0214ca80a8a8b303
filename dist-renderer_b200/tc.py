"""Tensor-core engine operand preparation (split-fp16 weight tiles for csrc/mlp_tc.cu)."""


def supported(plan):
    """True when the tcgen05 engine covers this decoder shape on this device."""
    return False


def prepare(plan):
    raise NotImplementedError("tensor-core engine not available yet")
