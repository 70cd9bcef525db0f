"""Import stub (test infrastructure): ITS/models/vmamba_layers.py:1 does `from mamba_ssm import Mamba`; the class is only
referenced by the dead MambaSS2D module and commented-out code, never by the ITS model."""


class Mamba:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("mamba_ssm is not installed in this image; the ITS model does not use it")
