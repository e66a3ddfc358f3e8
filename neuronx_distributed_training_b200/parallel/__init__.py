from . import state as parallel_state
from .state import initialize_model_parallel, destroy_model_parallel, init_distributed
from .layers import (ColumnParallelLinear, RowParallelLinear, GQAQKVColumnParallelLinear, ParallelEmbedding,
                     shard_master_weight, unshard_weights, param_is_tp_duplicate)
from .norm import RMSNorm, LayerNorm, get_norm
from .loss import parallel_cross_entropy, from_parallel_logits_to_logprobs, fused_linear_cross_entropy
from .mappings import (copy_to_tensor_model_parallel_region, reduce_from_tensor_model_parallel_region,
                       scatter_to_tensor_model_parallel_region, gather_from_tensor_model_parallel_region,
                       scatter_to_sequence_parallel_region, gather_from_sequence_parallel_region,
                       reduce_scatter_to_sequence_parallel_region, split_tensor_along_last_dim)
from .random import get_rng_tracker, model_parallel_manual_seed, set_random_seed
