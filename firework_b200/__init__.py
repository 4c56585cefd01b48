"""firework_b200 — B200-native implementation of firework's rendering hot path."""
from .api import *  # noqa: F401,F403  (the mirrored firework API)
