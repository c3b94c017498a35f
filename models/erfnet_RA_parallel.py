"""Drop-in location of the reference's module: ``from models.erfnet_RA_parallel import Net as Net_RAP``
(train_RAPFT_step1.py:33, train_new_task_step2.py:33, train_new_task_step3.py:33)."""
from mdil_ss_b200.erfnet_RA_parallel import *  # noqa: F401,F403
from mdil_ss_b200.erfnet_RA_parallel import (DownsamplerBlock, non_bottleneck_1d, non_bottleneck_1d_RAP, Encoder,  # noqa: F401
                                             UpsamplerBlock, Decoder, Net)
