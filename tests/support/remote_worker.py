"""TEST INFRASTRUCTURE: a worker in its own process.  Connects its slot state machines to the channel manager at
argv[1], serves LzySlotsApi + LzybPeerApi on a free port, produces `argv[3]` bytes of the counter pattern (seed
argv[4]) into the channel argv[2] as slot "/remote/out", prints `READY <host:port>` and waits for a line on stdin."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))

import oracle  # noqa: E402
from lzy_b200 import _host as H  # noqa: E402
from lzy_b200.grpc_api import remote_cm, serve  # noqa: E402
from tests.support import hostmem  # noqa: E402


def main() -> int:
    cm_target, channel, nbytes, seed = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    hostmem.install()
    cm = remote_cm.connect(cm_target)
    probe, port = serve("127.0.0.1:0")
    probe.stop(0)
    url = f"127.0.0.1:{port}"
    slots = H.Slots(cm, 0, url)
    server, _ = serve(url, slots=slots)
    ctx = slots.context("exec-mp", "producer")
    ctx.add_slot("/remote/out", channel, H.DIR_OUTPUT)
    ctx.before_execution()
    h = ctx.output_open("/remote/out", nbytes)
    hostmem.write(h, oracle.pattern(nbytes, seed).tobytes())
    ctx.output_done("/remote/out")
    ctx.after_execution(20000)
    print(f"READY {url}", flush=True)
    sys.stdin.readline()
    ctx.close()
    server.stop(0)
    slots.close()
    remote_cm.disconnect_all()
    hostmem.uninstall()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
