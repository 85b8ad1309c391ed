"""TEST INFRASTRUCTURE ONLY - see package docstring."""


class TensorBoardLogger:
    pass


class WandbLogger:
    pass
