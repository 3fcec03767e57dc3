"""Drop-in for the reference's ``pytorch/nn`` package: the MLP LLR estimators (nn/llr.py),
inference on the B200-native tensor-core kernel (csrc/mlp.cu)."""
