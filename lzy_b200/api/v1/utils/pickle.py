"""The request travels on the command line as base64(cloudpickle) — same encoding as
pylzy/lzy/api/v1/utils/pickle.py:7-12, so a request built by a stock pylzy client unpickles here."""
import base64
from typing import Any

import cloudpickle


def pickle(obj: Any) -> str:
    return base64.b64encode(cloudpickle.dumps(obj)).decode("ascii")


def unpickle(base64_str: str) -> Any:
    return cloudpickle.loads(base64.b64decode(base64_str.encode("ascii")))
