"""Import-only stand-in for loguru (used by the reference for logging only)."""
import logging

logger = logging.getLogger("ganslate-oracle")
