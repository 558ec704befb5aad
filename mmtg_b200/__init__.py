"""mmtg_b200 — B200-native (sm_100a) implementation of the MMTG training/generation hot path.

Drop-in surface (mirrors /root/reference/src): `MMTG` (src/model.py:330), `MyLoss` (src/loss.py:39),
`sample_sequence` / `top_k_top_p_filtering` (src/generate.py:64,97), `model_cfgs` / `data_config`
(src/configs.py:14,43). Everything runs in libmmtg_b200.so (hand-written CUDA); importing the
package does not load the library, the first call does and raises if it is missing.
"""
__version__ = "0.2.0"

from .configs import data_config, model_cfgs  # noqa: F401
from .generate import generate_samples, sample_sequence, sample_sequence_batch, top_k_top_p_filtering  # noqa: F401
from .loss import MyLoss  # noqa: F401
from .model import MMTG  # noqa: F401
from .optim import FusedAdamW  # noqa: F401

__all__ = ["MMTG", "MyLoss", "sample_sequence", "sample_sequence_batch", "top_k_top_p_filtering",
           "generate_samples", "model_cfgs", "data_config", "FusedAdamW"]
