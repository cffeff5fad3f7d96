"""VoiceFilter baseline (ReLU activations) - drop-in for the reference class of the same import
path (/root/reference/models/voicefilter/model.py:11), running on the B200 engine."""
from voicesplit_b200.module import MaskEstimator


class VoiceFilter(MaskEstimator):
    ACTIVATION = "relu"
