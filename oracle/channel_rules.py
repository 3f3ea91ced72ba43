"""oracle/channel_rules.py — TEST INFRASTRUCTURE.  Compact Python model of the reference channel
manager's matching rules, written from the Java (not from the C++ under test):

  bind / createProducer / createConsumer   lzy/channel-manager/.../services/SlotsService.java:67-145,440-519
  unbind                                    :148-188
  transferFailed                            :191-291
  transferCompleted                         :294-358
  getOrCreate                               lzy/channel-manager/.../services/ChannelService.java:39-116
  priorities / FIND_PRODUCER                .../db/PeerDao.java:44-53, .../db/PeerDaoImpl.java:61-65
  cascade on peer drop                      .../resources/db/channel-manager/migrations/V1_0__Init_tables.sql

Random tie-breaks are exposed as *sets* of admissible answers so a differential test can accept any
of them.  Idempotency replays are not modelled (the conformance tests cover them directly)."""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Set, Tuple

PRIMARY, BACKUP = 7, 5
PENDING, ACTIVE, COMPLETED, FAILED = "PENDING", "ACTIVE", "COMPLETED", "FAILED"


class Status(Exception):
    def __init__(self, code: str, msg: str = ""):
        super().__init__(f"{code}: {msg}")
        self.code = code


@dataclass
class Peer:
    id: str
    role: str          # "PRODUCER" | "CONSUMER"
    storage: bool
    url: str
    priority: int
    connected: bool


@dataclass
class Transfer:
    id: str
    frm: str
    to: str
    state: str


@dataclass
class Channel:
    id: str
    user: str
    execution: str
    producer_uri: Optional[str]
    consumer_uri: Optional[str]
    peers: Dict[str, Peer] = field(default_factory=dict)
    transfers: Dict[str, Transfer] = field(default_factory=dict)


class ChannelRules:
    def __init__(self):
        self.channels: Dict[str, Channel] = {}
        self.aborted: Dict[str, str] = {}
        self._ids = itertools.count(1)
        self.start_transfers: List[Tuple[str, str, str]] = []  # (channel, consumer id, producer id)

    def get_or_create(self, user, execution, producer_uri=None, consumer_uri=None) -> str:
        if producer_uri is None and consumer_uri is None:
            raise Status("INVALID_ARGUMENT")
        for c in self.channels.values():
            if (c.user, c.execution, c.producer_uri, c.consumer_uri) == (user, execution, producer_uri, consumer_uri):
                return c.id
        c = Channel(f"ch{next(self._ids)}", user, execution, producer_uri, consumer_uri)
        role = "PRODUCER" if producer_uri is not None else "CONSUMER"
        pid = f"storage-peer-{next(self._ids)}"
        c.peers[pid] = Peer(pid, role, True, producer_uri or consumer_uri, BACKUP, False)
        self.channels[c.id] = c
        return c.id

    def _chan(self, cid) -> Channel:
        if cid not in self.channels:
            raise Status("INTERNAL", "Cannot get channel")
        return self.channels[cid]

    @staticmethod
    def producer_candidates(c: Channel) -> Set[str]:
        prods = [p for p in c.peers.values() if p.role == "PRODUCER" and p.priority >= 0]
        if not prods:
            return set()
        best = max(p.priority for p in prods)
        return {p.id for p in prods if p.priority == best}

    def _drop_peer(self, c: Channel, pid: str):
        c.peers.pop(pid, None)
        for tid in [t for t, tr in c.transfers.items() if pid in (tr.frm, tr.to)]:
            del c.transfers[tid]

    def _new_transfer(self, c, frm, to, state) -> str:
        tid = f"t{next(self._ids)}"
        c.transfers[tid] = Transfer(tid, frm, to, state)
        return tid

    def bind_consumer(self, cid, peer_id, url, chosen_producer: Optional[str]):
        """`chosen_producer` is what the implementation under test answered; it must be admissible."""
        c = self._chan(cid)
        if peer_id in c.peers:
            raise Status("INTERNAL", "already associated")
        cands = self.producer_candidates(c)
        if not cands:
            assert chosen_producer is None
            c.peers[peer_id] = Peer(peer_id, "CONSUMER", False, url, PRIMARY, False)
            return None
        assert chosen_producer in cands, (chosen_producer, cands)
        c.peers[peer_id] = Peer(peer_id, "CONSUMER", False, url, PRIMARY, True)
        return self._new_transfer(c, chosen_producer, peer_id, ACTIVE)

    def bind_producer(self, cid, peer_id, url):
        """returns (storage consumer id or None, [(consumer id, transfer state)])"""
        c = self._chan(cid)
        if peer_id in c.peers:
            raise Status("INTERNAL", "already associated")
        c.peers[peer_id] = Peer(peer_id, "PRODUCER", False, url, PRIMARY, False)
        claimed = [p for p in c.peers.values() if p.role == "CONSUMER" and not p.connected]
        out, storage = [], None
        for p in claimed:
            p.connected = True
            st = ACTIVE if p.storage else PENDING
            self._new_transfer(c, peer_id, p.id, st)
            out.append((p.id, st))
            if p.storage and storage is None:
                storage = p.id
            if not p.storage:
                self.start_transfers.append((cid, p.id, peer_id))
        return storage, out

    def unbind(self, cid, peer_id):
        c = self._chan(cid)
        if any(peer_id in (t.frm, t.to) and t.state in (PENDING, ACTIVE) for t in c.transfers.values()):
            raise Status("UNAVAILABLE")
        self._drop_peer(c, peer_id)

    def transfer_completed(self, cid, tid):
        c = self._chan(cid)
        if tid not in c.transfers:
            raise Status("NOT_FOUND")
        t = c.transfers[tid]
        if t.state != ACTIVE:
            raise Status("FAILED_PRECONDITION")
        t.state = COMPLETED
        to = c.peers.get(t.to)
        if to is not None and to.storage:
            self._drop_peer(c, to.id)
            c.peers[to.id] = Peer(to.id, "PRODUCER", True, to.url, BACKUP, False)

    def transfer_failed(self, cid, tid, chosen_producer: Optional[str]):
        c = self._chan(cid)
        if tid not in c.transfers:
            raise Status("NOT_FOUND")
        t = c.transfers[tid]
        if t.state != ACTIVE:
            raise Status("FAILED_PRECONDITION")
        t.state = FAILED
        to = c.peers.get(t.to)
        if to is not None and to.storage:
            self.aborted[c.execution] = "Uploading data to storage failed"
            del self.channels[cid]
            raise Status("INTERNAL", "Uploading data to storage failed")
        if t.frm in c.peers:
            c.peers[t.frm].priority -= 1
        cands = self.producer_candidates(c)
        if not cands:
            self.aborted[c.execution] = "No more producers in channel, while there are consumers"
            del self.channels[cid]
            raise Status("INTERNAL", "No more producers")
        assert chosen_producer in cands, (chosen_producer, cands)
        return self._new_transfer(c, chosen_producer, t.to, ACTIVE)

    def mark_started(self, cid, consumer_id, producer_id):
        c = self.channels.get(cid)
        if c is None:
            return
        for t in c.transfers.values():
            if (t.frm, t.to, t.state) == (producer_id, consumer_id, PENDING):
                t.state = ACTIVE
                return

    def summary(self, cid):
        c = self.channels[cid]
        prods = sorted((p.id, p.priority, p.storage) for p in c.peers.values() if p.role == "PRODUCER")
        cons = sorted((p.id, p.connected, p.storage) for p in c.peers.values() if p.role == "CONSUMER")
        trs = sorted((t.frm, t.to, t.state) for t in c.transfers.values())
        return prods, cons, trs
