"""ssrlcv_b200 -- B200-native (sm_100a) SIFT descriptor matching behind the MatchFactory API of uga-ssrl/SSRLCV."""
from . import types  # noqa: F401

__all__ = ["types", "checkpoint", "unity", "match_factory", "pair_scheduler", "fusion"]
