"""Drop-in replacement of upstream ``src/losses`` (``__init__.py:5-11``):

    from motionpriorcmax_b200.losses import LossFactory
    loss_calculator = LossFactory.get_loss_calculator('FOCUS', loss_config)
"""
from .base import TrajectoryLossBase
from .focus import FocusLoss


class LossFactory:
    @staticmethod
    def get_loss_calculator(loss_name, loss_config, profiler=None):
        if loss_name == 'FOCUS':
            return FocusLoss(**loss_config, profiler=profiler)
        else:
            raise ValueError("Unsupported loss type")


__all__ = ["LossFactory", "TrajectoryLossBase", "FocusLoss"]
