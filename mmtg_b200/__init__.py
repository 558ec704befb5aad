"""mmtg_b200 — B200-native (sm_100a) implementation of the MMTG training/generation hot path.

Drop-in surface (mirrors /root/reference/src): `MMTG`, `MyLoss`, `sample_sequence`,
`top_k_top_p_filtering`, `model_cfgs`, `data_config`.
"""
__version__ = "0.1.0"
