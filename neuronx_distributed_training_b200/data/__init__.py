from .base import (BaseDataModule, SyntheticDataModule, HFDataModule, ResumableBatchSampler, SyntheticTokenDataset,
                   get_batch_on_this_context_parallel_rank, default_collate)
