"""The kept gRPC surface served FROM HBM ("no green -m gpu test serves Read from HBM").
A slots server over real sockets in front of the CUDA data plane: a *legacy* CPU consumer — a stock Java worker, here the
reference-shaped client — pulls an output slot that lives in an HBM frame with `LzySlotsApi.Read`: ≤ 2 MiB chunks, then
the EOS control message, the offset honoured (OutputSlot.ReadThread, lzy/slots/.../OutputSlot.java:238-265), bytes equal to
what the op wrote and to the oracle's chunked-stream restatement; status codes as the reference maps them
(lzy/slots/.../SlotsService.java:36,44,55,63)."""
import grpc
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from lzy_b200 import _host as H  # noqa: E402
from lzy_b200 import _native as N  # noqa: E402
from lzy_b200.grpc_api import ChannelManagerPrivateClient, SlotsApiClient, messages as M, serve  # noqa: E402


@pytest.fixture
def stack(native):
    cm = H.ChannelManager(seed=13)
    H.pump_start(cm)
    slots = H.Slots(cm, 0, "lzyb://grpc-gpu-worker")
    server, port = serve("127.0.0.1:0", slots=slots, cm=cm)
    target = f"127.0.0.1:{port}"
    priv, sapi = ChannelManagerPrivateClient(target), SlotsApiClient(target)
    yield cm, slots, priv, sapi
    priv.close()
    sapi.close()
    server.stop(0)
    slots.close()
    H.pump_stop(cm)
    cm.close()


def test_legacy_consumer_reads_an_hbm_frame_over_grpc(stack):
    cm, slots, priv, sapi = stack
    n = (5 << 20) + 12345
    a = np.random.default_rng(99).integers(0, 256, n, dtype=np.uint8)
    t = torch.from_numpy(a).cuda()
    sp = M.PeerDescription.StoragePeer(storage_uri="hbm://grpc-gpu/legacy")
    chan = priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="e", workflow_name="wf", user_id="u", consumer=sp)).channel_id
    ctx = slots.context("e", "task")
    ctx.add_slot("/task/out", chan, H.DIR_OUTPUT)
    ctx.before_execution()
    out = N.SlotHandle.adopt(ctx.output_open("/task/out", n), owned=False)
    out.set_meta(N.DT_U8, (n,))
    out.pack(t.data_ptr(), n)       # the CUDA pack kernel: frame + digest in HBM
    ctx.output_done("/task/out")
    ctx.after_execution(10000)
    try:
        sizes, got = [], bytearray()
        for chunk in sapi.call("Read", M.ReadDataRequest(peer_id="/task/out", offset=0)):
            if chunk.WhichOneof("kind") == "control":
                assert chunk.control == 0  # EOS
                break
            sizes.append(len(chunk.chunk))
            got += chunk.chunk
        assert sizes == [2 << 20, 2 << 20, (1 << 20) + 12345]
        ref = np.zeros(n, dtype=np.uint8)
        chunks, delivered = oracle.stream_copy(a, ref)  # OutputSlot.ReadThread → SlotInputTransfer, restated
        assert (chunks, delivered) == (len(sizes), n) and bytes(got) == ref.tobytes() == a.tobytes()
        off = (1 << 20) + 7
        assert sapi.read_all("/task/out", offset=off) == a.tobytes()[off:]
        assert sapi.read_all("/task/out", offset=n) == b""
        with pytest.raises(grpc.RpcError) as e:
            sapi.read_all("/task/missing")
        assert e.value.code() == grpc.StatusCode.NOT_FOUND
        assert N.stats()["d2h_bytes"] >= n  # the bytes came out of the frame, through the pinned bounce path of lzyb_slot_read
    finally:
        ctx.close()
        H.frame_store_drop("hbm://grpc-gpu/legacy")


def test_read_on_an_input_slot_is_unimplemented_over_grpc(stack):
    cm, slots, priv, sapi = stack
    sp = M.PeerDescription.StoragePeer(storage_uri="hbm://grpc-gpu/never")
    chan = priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="e2", workflow_name="wf", user_id="u", consumer=sp)).channel_id
    ctx = slots.context("e2", "cons")
    ctx.add_slot("/cons/in", chan, H.DIR_INPUT)
    ctx.prefetch()  # binds as an early consumer: registered, waiting for a producer
    try:
        with pytest.raises(grpc.RpcError) as e:
            sapi.read_all("/cons/in")
        assert e.value.code() == grpc.StatusCode.UNIMPLEMENTED  # InputSlot.java:99-102
    finally:
        ctx.close()
