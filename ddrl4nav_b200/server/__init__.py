from .utils import random_choice_prob_index, select_action
from .forward import forward_compute, ForwardModule
from .backward import BackwardModule
from .threads import (ForwardThread, BackwardQueue, BackwardThread, BackwardGetDataThread, BackwardTrainThread,
                      encode_forward_replies)

__all__ = ["random_choice_prob_index", "select_action", "forward_compute", "ForwardModule", "BackwardModule",
           "ForwardThread", "BackwardQueue", "BackwardThread", "BackwardGetDataThread", "BackwardTrainThread",
           "encode_forward_replies"]
