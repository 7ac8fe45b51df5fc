"""conch.platforms -> conch_b200.platform (reference conch/platforms/__init__.py)."""
from conch_b200.platform import current_platform  # noqa: F401
