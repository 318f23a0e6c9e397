from .utils import random_choice_prob_index, select_action
from .forward import forward_compute, ForwardModule
from .backward import BackwardModule

__all__ = ["random_choice_prob_index", "select_action", "forward_compute", "ForwardModule", "BackwardModule"]
