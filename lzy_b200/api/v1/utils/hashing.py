"""``md5_of_str`` is kept for result-entry hashes, whose input is a URI *string*
(pylzy/lzy/api/v1/utils/hashing.py:6-7; pylzy/lzy/core/workflow.py:266,280,293).  Payload hashing is
done on the GPU (include/lzyb_digest.h) — ``HashingIO`` has no counterpart here."""
import hashlib


def md5_of_str(uri: str) -> str:
    return hashlib.md5(uri.encode("utf-8")).hexdigest()
