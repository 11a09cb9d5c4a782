from .aggregation import DelayedAggregation
from .base import HyperParameter, Pipeline, PipelineConfig
from .clustering import OnlineSpeakerClustering
from .diarization import SpeakerDiarization, SpeakerDiarizationConfig
from .embedding import (EmbeddingNormalization, OverlapAwareSpeakerEmbedding, OverlappedSpeechPenalty,
                        SpeakerEmbedding)
from .segmentation import SpeakerSegmentation
from .utils import AdjustVolume, Binarize, Resample
from .vad import VoiceActivityDetection, VoiceActivityDetectionConfig
