from .clip_official import ClipModel
from .losses import MaskedContrastiveLoss
from .projections import MLPLayers
from .retrieval import mutualRetrieval
from .speech_encoder_plus import FairseqSpeechEncoder_Hubert
from .weighted_sum import WeightedSumLayer
