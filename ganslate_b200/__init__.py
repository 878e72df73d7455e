"""ganslate_b200 -- B200-native (sm_100a) implementation of ganslate's training hot path.

Drop-in for the reference's config-registered module API on that path only:
`ganslate_b200.nn.generators`, `ganslate_b200.nn.discriminators`, `ganslate_b200.nn.losses`,
`ganslate_b200.nn.gans.{paired,unpaired}` mirror `ganslate.nn.*` (same class names, constructor arguments,
state_dict keys).  All compute goes through the C ABI in include/ganslate_b200.h.
"""
__version__ = "0.1.0"
