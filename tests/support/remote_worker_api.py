"""TEST INFRASTRUCTURE: `python -m lzy_b200.grpc_api worker` with the in-memory backend instead of a GPU — one worker
process serving WorkerApi / LongRunningService / LzySlotsApi / LzybPeerApi on argv[2], its slot state machines
connected to the channel manager at argv[1] — over gRPC, or, when argv[3] names the channel manager's unix socket, over the
box-local transport (csrc/host/local_rpc.cpp); argv[4], if given, is the unix socket this worker answers Describe on.
Prints `READY` and runs until a line arrives on stdin."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))

from lzy_b200.grpc_api import serve  # noqa: E402
from lzy_b200.grpc_api.__main__ import close_worker, make_worker  # noqa: E402
from tests.support import hostmem  # noqa: E402


def main() -> int:
    cm_target, port = sys.argv[1], int(sys.argv[2])
    cm_local = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else None
    own_socket = sys.argv[4] if len(sys.argv) > 4 else None
    hostmem.install()
    cm, worker, _url = make_worker(0, "127.0.0.1", port, cm_target, cm_local, own_socket, mount_factory=hostmem.HostmemMount)
    server, _ = serve(f"127.0.0.1:{port}", worker=worker)
    print("READY", flush=True)
    sys.stdin.readline()
    server.stop(0)
    close_worker(cm, worker)
    hostmem.uninstall()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
