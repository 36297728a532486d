"""CPU test: `timer` behaves as the reference's tests expect (tests/test_utils.py in iscc-search)."""

import time
from io import StringIO

from iscc_search_b200 import timer


def test_timer_logs_start_and_completion_with_elapsed_seconds():
    from loguru import logger

    out = StringIO()
    sink = logger.add(out, format="{message}")
    with timer("Test operation"):
        time.sleep(0.01)
    with timer("Operation with start log", log_start=True):
        pass
    logger.remove(sink)
    text = out.getvalue()
    assert "Test operation - completed" in text and "seconds)" in text
    assert "Test operation - started" not in text
    assert float(text.split("(")[1].split(" ")[0]) >= 0.01
    assert "Operation with start log - started" in text and "Operation with start log - completed" in text
