from .gae import accumulate_rewards, accumulate_tempo_rewards, GAE

__all__ = ["accumulate_rewards", "accumulate_tempo_rewards", "GAE"]
