"""Format helpers on either side of ``SynthesizerTrn.infer``: bringing speech-encoder features to
the frame rate and extracting the per-frame volume.  Same names, arguments and results as the reference's
``utils.repeat_expand_2d`` (utils.py:396-424) and ``utils.Volume_Extractor`` (utils.py:560-572); the only difference is
how the 'left' expansion is executed: the reference assigns one output column per Python iteration (``target_len`` tiny
device kernels), here the column map is computed on the host with the reference's float32 comparisons and applied as ONE
gather."""
from __future__ import annotations

import numpy as np
import torch
from torch.nn import functional as F


def _left_index(src_len: int, target_len: int) -> np.ndarray:
    """Source column of every target column under the reference's sequential rule (utils.py:408-415): the cursor advances
    by at most one per output column, when ``i >= temp[cursor+1]`` with ``temp = arange(src_len+1)*target_len/src_len``
    evaluated in float32."""
    temp = (torch.arange(src_len + 1) * target_len / src_len).numpy()        # float32, like the reference
    idx = np.empty(target_len, dtype=np.int64)
    cur = 0
    for i in range(target_len):
        if not (np.float32(i) < temp[cur + 1]):
            cur += 1
        idx[i] = cur
    return idx


def repeat_expand_2d(content: torch.Tensor, target_len: int, mode: str = "left") -> torch.Tensor:
    """content [h, t] -> [h, target_len] (utils.py:396-424)."""
    if mode == "left":
        idx = torch.from_numpy(_left_index(content.shape[-1], target_len)).to(content.device)
        return content.to(torch.float).index_select(-1, idx)
    return F.interpolate(content[None, :, :], size=target_len, mode=mode)[0]


class Volume_Extractor:
    """Per-frame RMS of the audio (utils.py:560-572): reflect-pad by hop/2, mean of squares over each hop, sqrt."""

    def __init__(self, hop_size: int = 512):
        self.hop_size = hop_size

    def extract(self, audio):  # audio: [1, n] tensor / array
        if not isinstance(audio, torch.Tensor):
            audio = torch.Tensor(audio)
        n_frames = int(audio.size(-1) // self.hop_size)
        audio2 = audio ** 2
        audio2 = F.pad(audio2, (int(self.hop_size // 2), int((self.hop_size + 1) // 2)), mode="reflect")
        volume = F.unfold(audio2[:, None, None, :], (1, self.hop_size), stride=self.hop_size)[:, :, :n_frames].mean(dim=1)[0]
        return torch.sqrt(volume)


VolumeExtractor = Volume_Extractor
