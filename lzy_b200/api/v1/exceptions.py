class LzyExecutionException(Exception):
    """An op failed and its original exception could not be recovered
    (pylzy/lzy/api/v1/local/runtime.py:173)."""
