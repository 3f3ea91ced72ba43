"""``python -m lzy_b200.grpc_api channel-manager --port 18122`` — the channel manager as its own service:
``LzyChannelManager`` + ``LzyChannelManagerPrivate`` on one port, deferred actions delivered to the workers'
``LzySlotsApi`` over the wire (`RemoteActionPump`).  Needs no GPU: the matcher never touches payload bytes.
(The reference's counterpart is the `channel-manager` Micronaut service, lzy/channel-manager/src/main/java/ai/lzy/
channelmanager/ChannelManagerApp.java; its Postgres tables are the in-memory maps of csrc/host/channel_manager.cpp.)

``python -m lzy_b200.grpc_api worker --gpu 0 --port 18123 --channel-manager host:18122`` — one worker process pinned to
one GPU (WorkerApiImpl + Slots of the reference's worker VM, lzy/worker/src/main/java/ai/lzy/worker/Worker.java:36-56):
``WorkerApi``, ``LongRunningService``, ``LzySlotsApi`` and ``LzybPeerApi`` on one port; its slot state machines reach
the channel manager over the wire (``remote_cm``) and map producers that live in other worker processes
(``peers``).  Needs a CUDA device: there is no CPU data path.
"""
from __future__ import annotations

import argparse
import logging
import signal
import threading

from .. import _host as H
from .pump import RemoteActionPump
from .server import serve


def _worker(args) -> int:
    from .. import _native as N
    from ..worker import Worker
    from . import peers, remote_cm
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    try:
        N.init(args.gpu, 1, args.host_ring_mib << 20)
    except N.LzybError as e:
        print(f"worker needs a CUDA device: {e}", flush=True)
        return 3
    cm = remote_cm.connect(args.channel_manager)
    peers.install()
    url = f"cuda+ipc://{args.host}:{args.port}?gpu={args.gpu}"
    worker = Worker(cm, args.gpu, url)
    server, port = serve(f"{args.host}:{args.port}", worker=worker)
    print(f"worker for GPU {args.gpu} listening on {args.host}:{port} as {url}", flush=True)
    done = threading.Event()
    for sig in (signal.SIGINT, signal.SIGTERM):
        signal.signal(sig, lambda *_: done.set())
    done.wait()
    server.stop(1)
    worker.close()
    peers.uninstall()
    remote_cm.disconnect_all()
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m lzy_b200.grpc_api")
    sub = ap.add_subparsers(dest="service", required=True)
    cm_args = sub.add_parser("channel-manager", help="serve LzyChannelManager(+Private)")
    cm_args.add_argument("--host", default="127.0.0.1")
    cm_args.add_argument("--port", type=int, default=0)
    cm_args.add_argument("--seed", type=int, default=None, help="deterministic producer tie-break (tests)")
    w_args = sub.add_parser("worker", help="serve WorkerApi + LzySlotsApi + LzybPeerApi for one GPU")
    w_args.add_argument("--gpu", type=int, default=0)
    w_args.add_argument("--host", default="127.0.0.1")
    w_args.add_argument("--port", type=int, required=True, help="also part of the URL other workers dial: must be fixed")
    w_args.add_argument("--channel-manager", required=True, metavar="HOST:PORT")
    w_args.add_argument("--host-ring-mib", type=int, default=64)
    args = ap.parse_args(argv)
    if args.service == "worker":
        return _worker(args)

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
