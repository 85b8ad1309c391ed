"""B200-native contrast-maximisation (CMax) loss path of tub-rip/MotionPriorCMax.

Public surface (mirrors the upstream modules it replaces):
    motionpriorcmax_b200.losses.LossFactory / FocusLoss          <- src/losses
    motionpriorcmax_b200.utils.EventImageConverter               <- src/utils/event_image_converter.py
    motionpriorcmax_b200.trajectories                            <- src/utils/trajectories.py, basis.py
    motionpriorcmax_b200.synthetic                               <- shapes of src/loader/* batches
The arithmetic lives in csrc/*.cu behind the C ABI of include/cmax_b200.h.
"""
__version__ = "0.1.0"
