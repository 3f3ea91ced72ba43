"""User-visible helper types (pylzy/lzy/types.py): ``File`` — a path whose *content* travels."""
from __future__ import annotations

import pathlib


class File(type(pathlib.Path())):  # type: ignore[misc]
    """A pathlib path; when passed to / returned from an @op its bytes (and permissions) are the
    payload (pylzy/lzy/serialization/file.py:16-50)."""
