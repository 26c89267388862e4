"""pmcts_b200 -- B200-native rollout engine for IBoat-PMCTS.

Importing the package does not need a GPU; creating an :class:`engine.Engine` does.
"""
__version__ = "0.1.0"
