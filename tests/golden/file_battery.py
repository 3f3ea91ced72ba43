"""``File`` payloads through the ``raw_file`` serializer, twice: by tests/golden/make_golden.py with the *reference's own*
``FileSerializer`` (pylzy/lzy/serialization/file.py:16-50, loaded by file path in the authoring container, its serialzy
base class stubbed) to produce tests/golden/reference_file_vectors.json, and by tests/test_file_serializer.py with
lzy_b200.serialization.builtin.FileSerializer — the byte format (``permissions`` header + octal mode + content), and
what reading makes of streams with and without that header, shorter than the header, empty.

`ns`: ``File`` (that side's path class), ``serialize(file, dest)``, ``deserialize(source) -> File``."""
import os
from io import BytesIO
from typing import Any, Dict


def run(ns: Dict[str, Any], directory: str) -> Dict[str, Any]:
    out: Dict[str, Any] = {"written": {}, "read": {}}
    contents = {"empty": b"", "short": b"hi", "text": b"hello, world\n", "binary": bytes(range(256)) * 40,
                "looks-like-header": b"permissions0o644 and more"}
    for name, content in contents.items():
        for mode in (0o644, 0o755, 0o600):
            path = os.path.join(directory, f"{name}-{mode:o}")
            with open(path, "wb") as f:
                f.write(content)
            os.chmod(path, mode)
            dest = BytesIO()
            ns["serialize"](ns["File"](path), dest)
            out["written"][f"{name}/{mode:o}"] = dest.getvalue().hex() if len(content) < 64 else \
                {"head": dest.getvalue()[:32].hex(), "length": len(dest.getvalue())}

    streams = {"with header": b"permissions0o755payload", "header only": b"permissions0o600", "no header": b"just some bytes, no header",
               "shorter than the header": b"tiny", "exactly header length": b"elevenbytes", "empty": b"",
               "header then 4096+": b"permissions0o640" + b"x" * 5000}
    for name, raw in streams.items():
        f = ns["deserialize"](BytesIO(raw))
        with open(f, "rb") as h:
            content = h.read()
        out["read"][name] = {"content": content.hex() if len(content) < 64 else {"length": len(content), "head": content[:16].hex()},
                             "mode": oct(os.stat(f).st_mode & 0o777) if raw.startswith(b"permissions") else "default",
                             "is File": isinstance(f, ns["File"])}
        os.unlink(f)
    return out
