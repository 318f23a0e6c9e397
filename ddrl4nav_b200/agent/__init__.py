from .gae import accumulate_rewards, GAE

__all__ = ["accumulate_rewards", "GAE"]
