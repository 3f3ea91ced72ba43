"""TEST INFRASTRUCTURE: `python -m lzy_b200.grpc_api worker` with the in-memory backend instead of a GPU — one worker
process serving WorkerApi / LongRunningService / LzySlotsApi / LzybPeerApi on argv[2], its slot state machines
connected to the channel manager at argv[1].  Prints `READY` and runs until a line arrives on stdin."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))

from lzy_b200.grpc_api import peers, remote_cm, serve  # noqa: E402
from lzy_b200.worker import Worker  # noqa: E402
from tests.support import hostmem  # noqa: E402


def main() -> int:
    cm_target, port = sys.argv[1], int(sys.argv[2])
    hostmem.install()
    cm = remote_cm.connect(cm_target)
    peers.install()
    worker = Worker(cm, 0, f"cuda+ipc://127.0.0.1:{port}?gpu=0", mount_factory=hostmem.HostmemMount)
    server, _ = serve(f"127.0.0.1:{port}", worker=worker)
    print("READY", flush=True)
    sys.stdin.readline()
    server.stop(0)
    worker.close()
    peers.uninstall()
    remote_cm.disconnect_all()
    hostmem.uninstall()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
