"""
`timer` with the behaviour iscc-search re-exports from iscc_usearch (/root/reference/iscc_search/utils.py:3),
pinned by /root/reference/tests/test_utils.py: logs "<message> - started" (optional) and
"<message> - completed (X.XXXX seconds)" through loguru when available, else the stdlib logger.
"""

import time
from contextlib import contextmanager

try:  # loguru is what the reference logs with; optional here
    from loguru import logger as _logger
except ImportError:  # pragma: no cover
    import logging

    _logger = logging.getLogger("iscc_search_b200")


@contextmanager
def timer(message, log_start=False):
    # type: (str, bool) -> object
    """Context manager that logs the elapsed wall time of the enclosed block."""
    if log_start:
        _logger.info(f"{message} - started")
    start = time.perf_counter()
    try:
        yield
    finally:
        elapsed = time.perf_counter() - start
        _logger.info(f"{message} - completed ({elapsed:.4f} seconds)")
