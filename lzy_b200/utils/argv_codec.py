"""Objects that travel in an argv slot.

A task's request reaches the worker as the last token of its command line
(pylzy/lzy/api/v1/remote/runtime.py:375-381), so its encoding has to survive a shell word: cloudpickle bytes in
standard base64, ASCII only — the encoding of pylzy/lzy/api/v1/utils/pickle.py:7-12, byte for byte (golden vectors:
tests/golden/reference_pickle_vectors.json).  Besides the codec this module knows how to find the request in a
command line and refuses tokens that cannot be one."""
from __future__ import annotations

import binascii
from base64 import b64decode, b64encode
from typing import Any

import cloudpickle

_ALPHABET = frozenset(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/=")


def encode(obj: Any) -> str:
    return b64encode(cloudpickle.dumps(obj)).decode("ascii")


def decode(text: str) -> Any:
    raw = text.encode("ascii")
    if not _ALPHABET.issuperset(raw):
        raise ValueError("not a base64 token")
    try:
        blob = b64decode(raw, validate=True)
    except binascii.Error as e:
        raise ValueError(f"not a base64 token: {e}") from None
    return cloudpickle.loads(blob)


def request_token(command: str) -> str:
    """the encoded request of a startup command line: its last shell word"""
    words = command.split()
    if not words:
        raise ValueError("empty command")
    return words[-1]
