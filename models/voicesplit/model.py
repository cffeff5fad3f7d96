"""VoiceSplit mask estimator (Mish activations) - drop-in for the reference class of the same
import path (/root/reference/models/voicesplit/model.py:9), running on the B200 engine."""
from voicesplit_b200.module import MaskEstimator


class VoiceSplit(MaskEstimator):
    ACTIVATION = "mish"   # x * tanh(softplus(x)), reference utils/generic_utils.py:395-399
