from .trainer import Trainer, NLPTrainer
from .module import BaseModelModule, get_param_groups_by_weight_decay
