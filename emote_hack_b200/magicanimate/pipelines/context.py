"""Mirror of reference magicanimate/pipelines/context.py (sliding-window scheduler)."""
from ...pipeline import get_context_scheduler, uniform  # noqa: F401
from ...pipeline import _ordered_halving as ordered_halving  # noqa: F401
