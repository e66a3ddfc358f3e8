from .loader import (Config, ConfigList, ConfigError, load_config, load_yaml, from_dict, apply_overrides,
                     get_attribute_from_cfg, register_resolver)
from .schema import (PrecisionPolicy, ParallelDims, batch_arithmetic, padded_vocab_size, process_config,
                     NEURON_ONLY_KEYS)
