"""EMA-codebook variant of the indexing path: import surface of reference ``index_improve/`` (models, trainer, main).

Differences from ``index/``: the quantiser keeps EMA statistics per code and moves the codebook
towards the EMA means in every training step (``lcrec_ema_update``), resets dead codes every ``reset_interval`` steps and
reports codebook utilisation (``lcrec_codebook_usage``).  Encoder / decoder / Sinkhorn / argmin are the kernels of the
base package.  ``datasets.py`` and ``utils.py`` of the reference variant are identical to ``index/``'s and are re-exported.
"""
from ..datasets import EmbDataset  # noqa: F401
