"""Result object of the clustering step (the subset of reference ``src/diart/mapping.py`` that
``OnlineSpeakerClustering`` hands to its caller).

The cost-matrix algebra and the Hungarian re-solves of the reference's ``SpeakerMap`` run inside the
CUDA clustering kernel (``csrc/cluster.cu``); what comes back is the final local->global assignment.
``SpeakerMap`` here wraps that assignment as a *hard* map (reference ``mapping.py:22-46``): 0 on
assigned pairs, the invalid value 1e10 elsewhere, so ``valid_assignments``, ``to_dict``,
``is_source_speaker_mapped`` and ``apply`` (``mapping.py:214-243,341-360``) behave as on the
reference's final map.  The soft distances of intermediate maps are not materialised.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

INVALID_VALUE = 1e10


class SpeakerMap:
    def __init__(self, assignment: np.ndarray, num_target_speakers: int):
        self.assignment = np.asarray(assignment, dtype=np.int32)   # (K,), -1 = unmapped
        self.num_source_speakers = int(self.assignment.shape[0])
        self.num_target_speakers = int(num_target_speakers)

    @property
    def mapping_matrix(self) -> np.ndarray:
        m = np.ones((self.num_source_speakers, self.num_target_speakers)) * INVALID_VALUE
        for src, tgt in zip(*self.valid_assignments()):
            m[src, tgt] = 0.0
        return m

    @property
    def shape(self) -> Tuple[int, int]:
        return self.num_source_speakers, self.num_target_speakers

    @property
    def mapped_source_speakers(self) -> List[int]:
        return [int(s) for s in np.where(self.assignment >= 0)[0]]

    @property
    def mapped_target_speakers(self) -> List[int]:
        return sorted(int(t) for t in self.assignment if t >= 0)

    def __len__(self):
        return len(self.mapped_source_speakers)

    def valid_assignments(self, strict: bool = False, as_array: bool = False):
        src = self.mapped_source_speakers
        tgt = [int(self.assignment[s]) for s in src]
        if as_array:
            return np.array(src), np.array(tgt)
        return src, tgt

    def to_dict(self, strict: bool = False) -> Dict[int, int]:
        return dict(zip(*self.valid_assignments(strict)))

    def to_inverse_dict(self, strict: bool = False) -> Dict[int, int]:
        return {t: s for s, t in zip(*self.valid_assignments(strict))}

    def is_source_speaker_mapped(self, source_speaker: int) -> bool:
        return bool(self.assignment[source_speaker] >= 0)

    def is_target_speaker_mapped(self, target_speaker: int) -> bool:
        return target_speaker in self.mapped_target_speakers

    def apply(self, source_scores) -> np.ndarray:
        """(frames, local speakers) -> (frames, global speakers), float64, zeros where unknown."""
        data = source_scores.data if hasattr(source_scores, "sliding_window") else np.asarray(source_scores)
        projected = np.zeros((data.shape[0], self.num_target_speakers))
        for src, tgt in zip(*self.valid_assignments()):
            projected[:, tgt] = data[:, src]
        return projected
