"""``python -m lzy_b200.grpc_api channel-manager --port 18122`` — the channel manager as its own service:
``LzyChannelManager`` + ``LzyChannelManagerPrivate`` on one port, deferred actions delivered to the workers'
``LzySlotsApi`` over the wire (`RemoteActionPump`).  Needs no GPU: the matcher never touches payload bytes.
(The reference's counterpart is the `channel-manager` Micronaut service, lzy/channel-manager/src/main/java/ai/lzy/
channelmanager/ChannelManagerApp.java; its Postgres tables are the in-memory maps of csrc/host/channel_manager.cpp.)

Workers in this or other processes connect their slot state machines with ``remote_cm.connect("host:port")`` and
serve ``LzySlotsApi`` so the pump can reach them (tests/test_grpc_surface.py).
"""
from __future__ import annotations

import argparse
import logging
import signal
import threading

from .. import _host as H
from .pump import RemoteActionPump
from .server import serve


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m lzy_b200.grpc_api")
    sub = ap.add_subparsers(dest="service", required=True)
    cm_args = sub.add_parser("channel-manager", help="serve LzyChannelManager(+Private)")
    cm_args.add_argument("--host", default="127.0.0.1")
    cm_args.add_argument("--port", type=int, default=0)
    cm_args.add_argument("--seed", type=int, default=None, help="deterministic producer tie-break (tests)")
    args = ap.parse_args(argv)

    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    cm = H.ChannelManager(seed=args.seed) if args.seed is not None else H.ChannelManager()
    server, port = serve(f"{args.host}:{args.port}", cm=cm)
    pump = RemoteActionPump(cm, on_abort=lambda a: logging.getLogger("lzy_b200.cm").error(
        "workflow %s of execution %s must be aborted: %s", a["workflow_name"], a["execution_id"], a["reason"])).start()
    print(f"channel-manager listening on {args.host}:{port}", flush=True)
    done = threading.Event()
    for sig in (signal.SIGINT, signal.SIGTERM):
        signal.signal(sig, lambda *_: done.set())
    done.wait()
    pump.stop()
    server.stop(1)
    cm.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
