"""B200Runtime — the ``Runtime`` that replaces ``LocalRuntime`` / the remote worker stack for the
slot/channel data plane on one 8×B200 box.

What it stands in for (reference call stacks SURVEY §3.1/§3.2):

* ``LocalRuntime.exec`` / ``__exec_call`` (pylzy/lzy/api/v1/local/runtime.py:42-179): no temp dirs, no
  storage↔file copies, no ``python startup.py`` subprocess per op — ops run in-process on a worker
  pinned to a GPU, arguments are zero-copy views of delivered HBM frames;
* ``startup.process_execution`` (pylzy/lzy/api/v1/startup.py:109-170): same argument / result /
  exception semantics (type check of results, ``lazy_arguments``, exception tuple, ``-> None``);
* lzy-service ``CreateChannels`` / ``BuildTasks`` (lzy/lzy-service/.../operations/graph/
  CreateChannels.java:75-104, BuildTasks.java:72-126): one channel per data URI, created with a storage
  consumer for op outputs and a storage producer for pre-existing data; slot names
  ``/{call.id}/{entry.id}`` (pylzy/lzy/api/v1/remote/runtime.py:326-349);
* graph-executor dispatch (lzy/graph-executor/.../BfsGraphProcessor.java): the C++ GraphDispatcher;
* worker staging order (lzy/worker/.../WorkerApiImpl.java:191-225): beforeExecution → op → afterExecution
  through the C++ slots runtime, one ``Slots`` instance per GPU.

Downstream prefetch: every task's input slots bind as *early consumers* when the graph starts, so each
transfer begins the moment its producer binds — on the consumer GPU's copy stream, while other ops are
still computing — instead of when the whole upstream task has been reported successful.
"""
from __future__ import annotations

import os
import sys
import threading
import time
import uuid
from concurrent.futures import FIRST_COMPLETED, ThreadPoolExecutor, wait
from dataclasses import dataclass, field
from typing import TYPE_CHECKING, Any, Callable, Dict, List, Optional, Sequence, Tuple

from .... import _host as H
from .... import _native as N
from ....devmem import current_stream_ptr
from ....core.dataflow import plan_dataflow, sole_consumer_outputs
from ....env import applied_env
from ....proxy import proxy
from ....serialization.types import get_type
from ....utils import op_scope, tbpickle
from ..exceptions import LzyExecutionException
from ..runtime import ProgressStep, Runtime
from ..snapshot import FrameSnapshot, Snapshot
from ..utils.types import check_types_serialization_compatible, infer_real_types, is_subtype
from .store import FrameStore

if TYPE_CHECKING:
    from ....core.call import LzyCall
    from ....core.workflow import LzyWorkflow

_LAZY_ARG_ATTR = "__lzy_lazy_argument__"


@dataclass
class B200RuntimeConfig:
    gpus: Optional[Sequence[int]] = None        # CUDA ordinals to use; None = all visible
    host_ring_bytes: int = 64 << 20
    placement: str = "round_robin"               # "round_robin" | "locality" | "single"
    prefetch: bool = True
    max_parallel_ops: int = 0                    # 0 = 2 × GPUs
    republish_inputs: bool = False               # "<slot>-out" seeding (InputSlot.java:160-168); ops own their inputs when off
    fanout_threshold: int = 3                    # an entry consumed on ≥ this many *other* GPUs is delivered by one chunk-pipelined
                                                 # pull chain (lzyb_broadcast) instead of one unicast per consumer; 0 = never
    fanout_chunk_bytes: int = 0                  # 0 = device-chained kernel; > 0 = host-driven chain with that chunk size
    retain: str = "auto"                         # "auto": a frame is freed once no proxy / queued call can ask for it; "all": keep until clear()
    chunk_bytes: int = 0                         # 0 = one launch per transfer; the reference streams 2 MiB chunks
    peer_chunk_bytes: int = 128 << 20            # slices of a cross-GPU transfer (more than one GPU only): an op's pack launched next to a
                                                 # transfer gets the GPU at the next slice boundary instead of after the whole payload; 0 = off
    transfer_waves: int = 1                      # > 1: cross-GPU transfer kernels run as this many waves of CTAs and the ops' streams get high
                                                 # priority (tools/overlap_probe.py: pack 0.55 ms inside a pull that costs nothing extra alone).
                                                 # Off: two chains on two GPUs measured 28 ms with 16 waves against 21 ms with slices
    move_sole_consumer: bool = True              # an output with one consumer on the same GPU and no other reader is handed over, not copied
    lazy_commit: bool = True                     # an op's result is READY in stream order: its successor is not held up by the pack kernel
    copy_engine: str = "auto"                    # "auto" | "ldg" | "tma" | "dma" | "push"
    spill_dir: Optional[str] = None              # tier below HBM: evicted frames and cache=True results live here as files
    hbm_budget_bytes: int = 0                    # resident frame-store bytes above which LRU frames are evicted (0 = never)
    user_id: str = field(default_factory=lambda: os.environ.get("LZY_USER", "local-user"))

    # deployment knobs without touching code: LZY_B200_<FIELD> (upper case) overrides a field that was left at its
    # default; explicit keyword arguments always win.  Lists are comma separated, sizes accept k/m/g suffixes.
    ENV_PREFIX = "LZY_B200_"

    @classmethod
    def from_env(cls, environ: Optional[Dict[str, str]] = None, **explicit: Any) -> "B200RuntimeConfig":
        import dataclasses
        env = os.environ if environ is None else environ
        values: Dict[str, Any] = {}
        for f in dataclasses.fields(cls):
            raw = env.get(cls.ENV_PREFIX + f.name.upper())
            if raw is None or f.name in explicit:
                continue
            try:
                values[f.name] = cls._parse(f.name, raw)
            except ValueError as e:
                raise ValueError(f"{cls.ENV_PREFIX}{f.name.upper()}={raw!r}: {e}") from None
        values.update(explicit)
        return cls(**values)

    @staticmethod
    def _parse(name: str, raw: str) -> Any:
        raw = raw.strip()
        if name == "gpus":
            return None if raw.lower() in ("", "all") else [int(x) for x in raw.split(",") if x.strip()]
        if name in ("prefetch", "republish_inputs", "move_sole_consumer", "lazy_commit"):
            if raw.lower() in ("1", "true", "yes", "on"):
                return True
            if raw.lower() in ("0", "false", "no", "off"):
                return False
            raise ValueError("expected a boolean")
        if name in ("host_ring_bytes", "fanout_chunk_bytes", "chunk_bytes", "peer_chunk_bytes", "hbm_budget_bytes"):
            mult = {"k": 1 << 10, "m": 1 << 20, "g": 1 << 30}.get(raw[-1:].lower())
            return int(float(raw[:-1]) * mult) if mult else int(raw)
        if name in ("max_parallel_ops", "fanout_threshold", "transfer_waves"):
            return int(raw)
        if name == "spill_dir":
            return raw or None
        choices = {"placement": ("round_robin", "locality", "single"), "retain": ("auto", "all"),
                   "copy_engine": ("auto", "ldg", "tma", "dma", "push")}.get(name)
        if choices and raw not in choices:
            raise ValueError(f"expected one of {choices}")
        return raw


class _Fanout:
    """one output consumed on several other GPUs: delivered once per GPU by a pull chain, then every
    consumer takes a private same-GPU copy (the reference opens one Read stream per consumer,
    channel-manager/.../services/SlotsService.java:457-466)"""

    def __init__(self, eid: str, uri: str, src_gpu: Optional[int], gpus: List[int]):
        from concurrent.futures import Future
        self.eid, self.uri, self.src_gpu, self.gpus = eid, uri, src_gpu, gpus
        self.future: "Future[Dict[int, N.SlotHandle]]" = Future()
        self.owned: List[N.SlotHandle] = []
        self.started = False


class _OpFailed(Exception):
    def __init__(self, task: int, exc_info):
        super().__init__(f"op #{task} failed")
        self.task, self.exc_info = task, exc_info


class B200Runtime(Runtime):
    # start / exec / finish / abort never suspend when nobody runs a loop for them: LzyEventLoop drives them on the caller's
    # thread (exec then runs the graph right there instead of parking it on an executor)
    inline_coroutines = True
    _lazy_lock = threading.Lock()
    _lazy_users = 0

    def __init__(self, config: Optional[B200RuntimeConfig] = None, **kwargs):
        self.config = config or B200RuntimeConfig.from_env(**kwargs)
        # started workflows, innermost last: an op body may open a workflow of its own on this very runtime while
        # its caller's graph is still running (the reference gives it a process of its own)
        self._active: List[Tuple["LzyWorkflow", str]] = []
        self._lock = threading.RLock()
        self._ready = False
        self.gpus: List[int] = []
        self.store: Optional[FrameStore] = None
        self.cm: Optional[H.ChannelManager] = None
        self.workers: Dict[int, H.Slots] = {}
        self._streams: Dict[int, Any] = {}
        self._pool: Optional[ThreadPoolExecutor] = None
        self._rr = 0
        self.stats = {"ops": 0, "graphs": 0, "cached_ops": 0}
        self.trace: Optional[list] = [] if os.environ.get("LZYB_TRACE") else None  # (task, phase, ms, t) records
        tbpickle.install()

    # ---- bring-up ------------------------------------------------------------------------------------
    def _ensure_ready(self) -> None:
        with self._lock:
            if self._ready:
                return
            import torch
            if not torch.cuda.is_available():
                raise N.NoGpuError(N.E_NOGPU, "B200Runtime needs a CUDA device: the data plane has no CPU path")
            if self.config.gpus is None:
                n = N.init(0, 0, self.config.host_ring_bytes)
                self.gpus = list(range(n))
            else:
                first, last = min(self.config.gpus), max(self.config.gpus)
                N.init(first, last - first + 1, self.config.host_ring_bytes)
                self.gpus = list(self.config.gpus)
            N.set_copy_engine({"auto": N.ENGINE_AUTO, "ldg": N.ENGINE_LDG, "tma": N.ENGINE_TMA,
                               "dma": N.ENGINE_DMA, "push": N.ENGINE_PUSH}[self.config.copy_engine])
            if self.config.lazy_commit:  # a process-wide switch of the library: on while any runtime that wants it is up
                with B200Runtime._lazy_lock:
                    B200Runtime._lazy_users += 1
                    N.set_lazy_commit(True)
            self.store = FrameStore(default_gpu=self.gpus[0])
            if self.config.spill_dir or self.config.hbm_budget_bytes:
                self.store.configure_tier(self.config.spill_dir, self.config.hbm_budget_bytes)
            self.cm = H.ChannelManager()
            H.pump_start(self.cm)
            waves = self.config.transfer_waves if len(self.gpus) > 1 else 1
            N.set_transfer_waves(max(1, waves))
            for g in self.gpus:
                w = H.Slots(self.cm, g, f"lzyb://localhost/worker?gpu={g}")
                w.set_republish(self.config.republish_inputs)
                if self.config.chunk_bytes:
                    w.set_chunk_bytes(self.config.chunk_bytes)
                if self.config.peer_chunk_bytes and len(self.gpus) > 1:
                    w.set_peer_chunk_bytes(self.config.peer_chunk_bytes)
                self.workers[g] = w
                self._streams[g] = torch.cuda.Stream(device=g, priority=-1 if waves > 1 else 0)
            workers = self.config.max_parallel_ops or 2 * len(self.gpus)
            self._pool = ThreadPoolExecutor(max_workers=max(2, workers), thread_name_prefix="lzyb-op")
            self._ready = True

    # ---- Runtime ABC ----------------------------------------------------------------------------------
    async def storage(self) -> Optional[Any]:
        return None  # data never leaves the box: there is no external storage to register

    def snapshot(self, workflow: "LzyWorkflow", storage_uri_prefix: str) -> Optional[Snapshot]:
        self._ensure_ready()
        snap = FrameSnapshot(workflow.owner.serializer_registry, storage_uri_prefix, self.store, workflow.owner.storage_name)
        snap.auto_release = self.config.retain == "auto"
        return snap

    async def start(self, workflow: "LzyWorkflow") -> str:
        self._ensure_ready()
        execution_id = str(uuid.uuid4())
        with self._lock:
            self._active.append((workflow, execution_id))
        return execution_id

    def _stop(self, reason: str, workflow: Optional["LzyWorkflow"] = None) -> None:
        """abort() / finish(): of `workflow` if given (nothing to do when it never started here), else of the innermost
        started workflow; nothing to do when none was started (pylzy/lzy/api/v1/remote/runtime.py:204-243: `if not
        self.__running: return`, state reset afterwards)"""
        with self._lock:
            if workflow is None:
                if not self._active:
                    return
                _workflow, execution_id = self._active.pop()
            else:
                for i in range(len(self._active) - 1, -1, -1):
                    if self._active[i][0] is workflow:
                        _workflow, execution_id = self._active.pop(i)
                        break
                else:
                    return
        if self.cm is not None:
            self.cm.destroy_all(execution_id, reason)  # LzyChannelManagerPrivate.DestroyAll

    async def abort(self) -> None:
        self._stop("aborted")

    async def finish(self) -> None:
        self._stop("finished")

    # LzyWorkflow prefers these two when a runtime has them: with several executions alive (an op that opened a
    # workflow of its own) "the innermost one" is not always the one that is ending
    async def abort_workflow(self, workflow: "LzyWorkflow") -> None:
        self._stop("aborted", workflow)

    async def finish_workflow(self, workflow: "LzyWorkflow") -> None:
        self._stop("finished", workflow)

    def _execution_of(self, workflow: "LzyWorkflow") -> str:
        with self._lock:
            for wf, execution_id in reversed(self._active):
                if wf is workflow:
                    return execution_id
        raise ValueError("Runtime is not running")  # remote/runtime.py:146-147

    async def exec(self, calls: List["LzyCall"], progress: Callable[[ProgressStep], None]) -> None:
        import asyncio
        with self._lock:
            if not self._active:
                raise ValueError("Runtime is not running")
        progress(ProgressStep.EXECUTING)
        try:
            loop = asyncio.get_running_loop()
        except RuntimeError:
            loop = None  # driven inline by LzyEventLoop.run_async on the caller's thread
        try:
            if loop is None:
                self._exec_graph(list(calls))
            else:
                await loop.run_in_executor(None, self._exec_graph, list(calls))
        except BaseException:
            progress(ProgressStep.FAILED)
            raise
        progress(ProgressStep.COMPLETED)

    def transfer_log(self, clear: bool = False) -> List[dict]:
        """per-transfer records of every worker (bytes, ns, final peer, failovers, outcome), tagged with the
        consumer's GPU — include/lzy_b200_host.h: lzyb_slots_transfer_log"""
        out: List[dict] = []
        for gpu, worker in self.workers.items():
            for rec in worker.transfer_log(clear):
                rec["gpu"] = gpu
                out.append(rec)
        return out

    def clear(self) -> None:
        """drop every stored frame (results of finished workflows stay readable until then)"""
        if self.store is not None:
            self.store.clear()

    def close(self) -> None:
        with self._lock:
            if not self._ready:
                return
            self.clear()
            for w in self.workers.values():
                w.close()
            H.pump_stop(self.cm)
            self.cm.close()
            if self._pool:
                self._pool.shutdown(wait=False)
            self.workers.clear()
            if self.config.lazy_commit:
                with B200Runtime._lazy_lock:
                    B200Runtime._lazy_users -= 1
                    if B200Runtime._lazy_users <= 0:
                        B200Runtime._lazy_users = 0
                        N.set_lazy_commit(False)
            self._ready = False

    # ---- graph execution ----------------------------------------------------------------------------------
    def _place(self, calls: List["LzyCall"], producers: Dict[str, int], snapshot: Snapshot) -> List[int]:
        placement: List[int] = []
        for i, call in enumerate(calls):
            prov = call.get_provisioning()
            if prov.gpu_index is not None and prov.gpu_index in self.gpus:
                placement.append(prov.gpu_index)
                continue
            if self.config.placement == "single" or len(self.gpus) == 1:
                placement.append(self.gpus[0])
            elif self.config.placement == "locality":
                best, best_bytes = None, -1
                for eid in list(call.arg_entry_ids) + list(call.kwarg_entry_ids.values()):
                    if eid in producers and producers[eid] < i:
                        g, nb = placement[producers[eid]], 1
                    else:
                        uri = snapshot.get(eid).storage_uri
                        g, nb = self.store.gpu_of(uri), self.store.nbytes_of(uri)
                    if g is not None and nb > best_bytes:
                        best, best_bytes = g, nb
                placement.append(best if best is not None else self.gpus[self._rr % len(self.gpus)])
            else:
                placement.append(self.gpus[self._rr % len(self.gpus)])
            self._rr += 1
        return placement

    def _exec_graph(self, calls: List["LzyCall"]) -> None:
        if not calls:
            return
        wf = calls[0].workflow  # the graph's own workflow: not necessarily the innermost one that is active
        assert self.store is not None and self.cm is not None
        snapshot = wf.snapshot
        exec_id, user = self._execution_of(wf), self.config.user_id
        self.stats["graphs"] += 1

        # ops whose outputs are all already supplied — in storage, or by an earlier op of this graph — are dropped
        # (lzy-service DataFlowGraph.java:51-72); the graph is keyed by data URI
        plan = plan_dataflow(calls, lambda eid: snapshot.get(eid).storage_uri, self.store.exists)
        self.stats["cached_ops"] += len(plan.cached)
        self.stats["duplicate_ops"] = self.stats.get("duplicate_ops", 0) + len(plan.duplicates)
        if hasattr(snapshot, "unpin"):
            for c in plan.dropped:  # a dropped call releases its inputs too
                for eid in list(c.arg_entry_ids) + list(c.kwarg_entry_ids.values()):
                    snapshot.unpin(eid)
        if not plan.calls:
            return
        calls = plan.calls
        producers: Dict[str, int] = plan.supplier  # result entry -> the op that supplies its URI
        inputs_of: List[List[str]] = []
        edges: List[Tuple[int, int]] = []
        for i, c in enumerate(calls):
            ins = list(dict.fromkeys(list(c.arg_entry_ids) + list(c.kwarg_entry_ids.values())))
            inputs_of.append(ins)
            for eid in ins:
                if eid in producers and producers[eid] != i:
                    edges.append((producers[eid], i))

        # CreateChannels: one channel per data URI
        channel_of: Dict[str, str] = {}
        for i, c in enumerate(calls):
            for eid in c.entry_ids:
                channel_of[eid] = self.cm.get_or_create(user, wf.name, exec_id, consumer_uri=snapshot.get(eid).storage_uri)
        for eid in producers:  # results of dropped duplicates share their supplier's URI, hence its channel
            if eid not in channel_of:
                channel_of[eid] = self.cm.get_or_create(user, wf.name, exec_id, consumer_uri=snapshot.get(eid).storage_uri)
        for i, ins in enumerate(inputs_of):
            for eid in ins:
                if eid in channel_of:
                    continue
                uri = snapshot.get(eid).storage_uri
                if not self.store.exists(uri):
                    raise ValueError(f"Cannot find blob in storage by uri={uri}")
                channel_of[eid] = self.cm.get_or_create(user, wf.name, exec_id, producer_uri=uri)

        placement = self._place(calls, producers, snapshot)
        graph = H.GraphDispatcher(len(calls), edges, allow_cycles=False)
        ctxs: List[H.ExecutionContext] = []

        # fan-out plan: which entries are consumed on many other GPUs
        fanout: Dict[str, _Fanout] = {}
        if self.config.fanout_threshold and len(self.gpus) > self.config.fanout_threshold:
            where: Dict[str, set] = {}
            for i, ins in enumerate(inputs_of):
                for eid in ins:
                    where.setdefault(eid, set()).add(placement[i])
            for eid, gset in where.items():
                uri = snapshot.get(eid).storage_uri
                src_gpu = placement[producers[eid]] if eid in producers else self.store.gpu_of(uri)
                remote = sorted(g for g in gset if g != src_gpu)
                if len(remote) >= self.config.fanout_threshold:
                    fanout[eid] = _Fanout(eid, uri, src_gpu, remote)
        # outputs that travel by ownership hand-over: one consumer on the producer's GPU and nobody else who could ask
        sole: set = set()
        if self.config.move_sole_consumer and self.config.retain == "auto" and not self.config.republish_inputs and \
                hasattr(snapshot, "claims"):
            sole = sole_consumer_outputs(plan, inputs_of, placement, snapshot.claims,
                                         lambda eid: eid in fanout or snapshot.is_persistent(eid))
            self.stats["moved_outputs"] = self.stats.get("moved_outputs", 0) + len(sole)
        try:
            # BuildTasks: slot → channel assignments; every input binds right away (prefetch)
            for i, c in enumerate(calls):
                ctx = self.workers[placement[i]].context(exec_id, c.id)
                for eid in inputs_of[i]:
                    if eid in fanout:
                        continue  # delivered by the fan-out chain, not by a unicast transfer
                    ctx.add_slot(f"/{c.id}/{eid}", channel_of[eid], H.DIR_INPUT)
                for eid in c.entry_ids:
                    ctx.add_slot(f"/{c.id}/{eid}", channel_of[eid], H.DIR_OUTPUT, sole_consumer=eid in sole)
                ctxs.append(ctx)
            if self.config.prefetch:
                for ctx in ctxs:
                    ctx.prefetch()

            for f in fanout.values():  # data that exists already can start travelling now
                if f.eid not in producers:
                    f.started = True
                    self._pool.submit(self._do_fanout, f)
            running: Dict[Any, int] = {}
            failure: Optional[_OpFailed] = None
            while True:
                ready = list(graph.next())
                while len(ready) == 1 and not running and failure is None:
                    # a chain: nothing to run next to this op, so it runs right here — no pool hand-off, no future,
                    # no wake-up (≈ 0.1 ms per op that the data plane does not have to give)
                    t = ready[0]
                    try:
                        self._run_task(t, calls[t], ctxs[t], placement[t], inputs_of[t], fanout)
                    except _OpFailed as err:
                        graph.task_done(t, False, repr(err))
                        failure = err
                        ready = []
                        break
                    graph.task_done(t, True)
                    for eid in plan.outputs_of(t):
                        f = fanout.get(eid)
                        if f is not None and not f.started:
                            f.started = True
                            self._pool.submit(self._do_fanout, f)
                    ready = list(graph.next())
                if failure is not None:
                    graph.stop("op failed")
                    break
                for t in ready:
                    running[self._pool.submit(self._run_task, t, calls[t], ctxs[t], placement[t], inputs_of[t], fanout)] = t
                if not running:
                    break
                done, _ = wait(list(running), return_when=FIRST_COMPLETED)
                for fut in done:
                    t = running.pop(fut)
                    err = fut.exception()
                    if err is None:
                        graph.task_done(t, True)
                        for eid in plan.outputs_of(t):  # its outputs are bound: start their fan-out
                            f = fanout.get(eid)
                            if f is not None and not f.started:
                                f.started = True
                                self._pool.submit(self._do_fanout, f)
                    else:
                        graph.task_done(t, False, repr(err))
                        if failure is None:
                            failure = err if isinstance(err, _OpFailed) else _OpFailed(t, (type(err), err, err.__traceback__))
                if failure is not None:
                    graph.stop("op failed")  # first failure stops the graph (BfsGraphProcessor.java:64-68)
                    for fut in list(running):
                        fut.cancel()
                    wait(list(running))
                    break
            if failure is not None:
                _typ, value, tb = failure.exc_info
                raise value.with_traceback(tb)
            if graph.status != H.G_COMPLETED:
                raise LzyExecutionException("graph did not complete")
        finally:
            mark = self._phase_marker(-1)
            for f in fanout.values():
                if not f.future.done():
                    f.future.set_exception(LzyExecutionException("graph stopped before the fan-out ran"))
                for sl in f.owned:
                    sl.close()
            for ctx in ctxs:
                ctx.close()
            mark("contexts closed")
            graph.close()
            if hasattr(snapshot, "unpin"):
                for i in range(len(calls)):
                    for eid in plan.outputs_of(i):  # dropped duplicates hold the shared frame in their own right
                        snapshot.produced(eid)
                for c in calls:
                    for eid in list(c.arg_entry_ids) + list(c.kwarg_entry_ids.values()):
                        snapshot.unpin(eid)
            mark("entries released")

    def _do_fanout(self, f: _Fanout) -> None:
        try:
            src = self.store._slot(f.uri)
            if src is None:
                raise ValueError(f"Cannot find blob in storage by uri={f.uri}")
            dsts = [N.SlotHandle.open_input(f"fanout/{f.eid}/gpu{g}", g, src.nbytes) for g in f.gpus]
            f.owned.extend(dsts)
            N.broadcast(src, dsts, self.config.fanout_chunk_bytes)
            per_gpu = {g: d for g, d in zip(f.gpus, dsts)}
            per_gpu[src.gpu] = src
            self.stats["fanouts"] = self.stats.get("fanouts", 0) + 1
            f.future.set_result(per_gpu)
        except BaseException as e:  # the consumers re-raise it
            if not f.future.done():
                f.future.set_exception(e)

    # ---- one op (WorkerApiImpl.executeOp + startup.process_execution) -------------------------------------
    def _phase_marker(self, task: int) -> Callable[[str], None]:
        """LZYB_TRACE=1: (task, phase, ms since the previous mark, perf_counter at the mark) records in `self.trace`;
        otherwise a no-op"""
        if self.trace is None:
            return lambda phase: None
        last = [time.perf_counter()]

        def mark(phase: str) -> None:
            now = time.perf_counter()
            self.trace.append((task, phase, 1e3 * (now - last[0]), now))
            last[0] = now

        return mark

    def _run_task(self, t: int, call: "LzyCall", ctx: H.ExecutionContext, gpu: int, input_eids: List[str],
                  fanout: Optional[Dict[str, _Fanout]] = None) -> None:
        import torch
        wf = call.workflow
        assert self.store is not None
        snapshot, registry = wf.snapshot, wf.owner.serializer_registry
        exc_uri = snapshot.get(call.exception_id).storage_uri
        mark = self._phase_marker(t)
        try:
            ctx.before_execution(-1)  # every input fully delivered and verified
            mark("before_execution")
            values: Dict[str, Any] = {}

            def load(eid: str) -> Any:
                if eid not in values:
                    if fanout and eid in fanout:
                        # the chain put one copy on this GPU; the op gets a private same-GPU copy of it
                        base = fanout[eid].future.result()[gpu]
                        slot = N.SlotHandle.open_input(f"/{call.id}/{eid}", gpu, base.nbytes)
                        if base.nbytes:
                            slot.transfer_chunk(base, 0, base.nbytes)
                        slot.finish(base)
                    else:
                        slot = N.SlotHandle.adopt(ctx.input_handle(f"/{call.id}/{eid}"), owned=True)
                    # the op body runs on this worker's stream: that is where the view must be ordered behind its producer
                    values[eid] = self.store.unpack_slot(slot, snapshot.get(eid).typ, registry, private=True,
                                                         stream=int(self._streams[gpu].cuda_stream))
                return values[eid]

            def arg(eid: str) -> Any:
                if call.lazy_arguments:
                    return proxy(lambda e=eid: load(e), infer_real_types(snapshot.get(eid).typ), cls_attrs={_LAZY_ARG_ATTR: True})
                return load(eid)

            args = [arg(e) for e in call.arg_entry_ids]
            kwargs = {k: arg(e) for k, e in call.kwarg_entry_ids.items()}
            func = call.signature.func.callable
            stream = self._streams[gpu]
            # one device / stream scope for the body and the packing of its results (entering it costs tens of µs)
            with torch.cuda.device(gpu), torch.cuda.stream(stream):
                op_scope.enter()
                try:
                    with applied_env(call.get_env_vars()):
                        res = func(*args, **kwargs)
                finally:
                    op_scope.leave()
                mark("load+op")
                outs = call.entry_ids
                results = [res] if len(outs) == 1 else list(res)
                sp = current_stream_ptr(gpu)
                for eid, data in zip(outs, results):
                    entry = snapshot.get(eid)
                    if hasattr(type(data), _LAZY_ARG_ATTR):  # the op returned one of its lazy arguments
                        data = data.__lzy_origin__
                    real_type = get_type(data)
                    compatible = check_types_serialization_compatible(entry.typ, real_type, registry)
                    if not compatible or not is_subtype(real_type, entry.typ):
                        raise TypeError(f"Invalid types: return value has type {entry.typ} but passed type {real_type}")
                    name = f"/{call.id}/{eid}"

                    def factory(nbytes, device, _name=name):
                        return N.SlotHandle.adopt(ctx.output_open(_name, nbytes), owned=False)

                    slot = self.store.pack_object(data, entry.typ, registry, factory, sp)
                    ctx.output_done(name)
                    values[f"out:{eid}"] = slot  # keep the packed source alive until the commit
            mark("pack")
            ctx.after_execution(-1)  # outputs bound; "uploaded" = referenced by the HBM frame store
            mark("after_execution")
            self.store.put_inline(exc_uri, None)
            self.stats["ops"] += 1
        except BaseException:
            info = sys.exc_info()
            try:
                self.store.put_inline(exc_uri, info)  # pickled (traceback included) and framed into the host ring
            except Exception:
                self.store.put_inline(exc_uri, (info[0], info[1], None))
            raise _OpFailed(t, info)
