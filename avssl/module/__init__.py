from .clip_official import ClipModel
from .losses import MaskedContrastiveLoss, SupConLoss
from .pooling import AttentivePoolingLayer, MeanPoolingLayer
from .projections import *
from .retrieval import mutualRetrieval
from .speech_encoder_plus import FairseqSpeechEncoder_Hubert, S3prlSpeechEncoderPlus
from .weighted_sum import WeightedSumLayer
