"""TEST INFRASTRUCTURE: `python -m lzy_b200.grpc_api worker …` itself — argument parsing, wiring, serving, signal handling,
teardown — on a box without a GPU: the in-memory backend is installed at the backend seam and `lzyb_init` is stood in for
(everything after it in `_worker` is the code a GPU worker runs).  Usage: the arguments of `python -m lzy_b200.grpc_api`."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))

from tests.support import hostmem  # noqa: E402

hostmem.install()
from lzy_b200 import _native as N  # noqa: E402

N.init = lambda *a, **k: 1  # no CUDA device here; the frames live in the in-memory backend

import lzy_b200.grpc_api.__main__ as service  # noqa: E402

_make_worker = service.make_worker
# ... and the op side mounts those frames instead of HBM frames (the one other thing of a GPU worker that needs the device)
service.make_worker = lambda *a, **k: _make_worker(*a, **{**k, "mount_factory": hostmem.HostmemMount})
main = service.main

if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))
