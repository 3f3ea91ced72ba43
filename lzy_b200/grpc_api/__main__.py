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


def make_worker(gpu: int, host: str, port: int, channel_manager: str, channel_manager_local=None, local_socket=None,
                mount_factory=None):
    """(matcher connection, Worker, the worker's URL) as the `worker` service wires them — one place, so that the CPU twin of
    the tests (tests/support/remote_worker_api.py, in-memory backend) runs the very wiring the GPU worker runs.
    The matcher: over gRPC, or — same box, this repository's channel manager — over its unix socket (no interpreter on the
    path of Bind / TransferCompleted; the matcher's deferred StartTransfer still arrives on the gRPC port)."""
    from ..worker import Worker
    from . import peers, remote_cm
    cm = H.LocalChannelManager(channel_manager_local) if channel_manager_local else remote_cm.connect(channel_manager)
    peers.install()
    url = f"cuda+ipc://{host}:{port}?gpu={gpu}"
    if local_socket:
        url += f"&uds={local_socket}"  # consumers of this box fetch this worker's frame descriptors there
    return cm, Worker(cm, gpu, url, local_socket=local_socket, mount_factory=mount_factory), url


def close_worker(cm, worker) -> None:
    from . import peers, remote_cm
    worker.close()
    peers.uninstall()
    if isinstance(cm, H.LocalChannelManager):
        cm.close()
    remote_cm.disconnect_all()


def _worker(args) -> int:
    from .. import _native as N
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    try:
        N.init(args.gpu, 1, args.host_ring_mib << 20)
    except N.LzybError as e:
        print(f"worker needs a CUDA device: {e}", flush=True)
        return 3
    cm, worker, url = make_worker(args.gpu, args.host, args.port, args.channel_manager, args.channel_manager_local, args.local_socket)
    server, port = serve(f"{args.host}:{args.port}", worker=worker)
    print(f"worker for GPU {args.gpu} listening on {args.host}:{port} as {url}", flush=True)
    done = threading.Event()
    for sig in (signal.SIGINT, signal.SIGTERM):
        signal.signal(sig, lambda *_: done.set())
    done.wait()
    server.stop(1)
    close_worker(cm, worker)
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m lzy_b200.grpc_api")
    sub = ap.add_subparsers(dest="service", required=True)
    cm_args = sub.add_parser("channel-manager", help="serve LzyChannelManager(+Private)")
    cm_args.add_argument("--host", default="127.0.0.1")
    cm_args.add_argument("--port", type=int, default=0)
    cm_args.add_argument("--seed", type=int, default=None, help="deterministic producer tie-break (tests)")
    cm_args.add_argument("--local-socket", default=None, metavar="PATH",
                         help="also serve Bind / Unbind / TransferCompleted / TransferFailed on this unix socket (workers of this box)")
    w_args = sub.add_parser("worker", help="serve WorkerApi + LzySlotsApi + LzybPeerApi for one GPU")
    w_args.add_argument("--gpu", type=int, default=0)
    w_args.add_argument("--host", default="127.0.0.1")
    w_args.add_argument("--port", type=int, required=True, help="also part of the URL other workers dial: must be fixed")
    w_args.add_argument("--channel-manager", required=True, metavar="HOST:PORT")
    w_args.add_argument("--host-ring-mib", type=int, default=64)
    w_args.add_argument("--channel-manager-local", default=None, metavar="PATH",
                        help="the channel manager's --local-socket: the slot state machines dial it instead of the gRPC port")
    w_args.add_argument("--local-socket", default=None, metavar="PATH",
                        help="answer LzybPeerApi.Describe on this unix socket too and say so in the worker's URL (uds=PATH)")
    args = ap.parse_args(argv)
    if args.service == "worker":
        return _worker(args)

    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    cm = H.ChannelManager(seed=args.seed) if args.seed is not None else H.ChannelManager()
    server, port = serve(f"{args.host}:{args.port}", cm=cm)
    if args.local_socket:
        cm.serve_local(args.local_socket)
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
