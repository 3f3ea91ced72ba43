"""The reference's gRPC surface for this path, served from the B200 data plane:
``LzySlotsApi`` (slots-api.proto:13-47), ``LzyChannelManager`` (channel-manager.proto:13-95) and
``LzyChannelManagerPrivate`` (channel-manager-private.proto:13-84), plus (§8f rank 4) ``WorkerApi``
(worker-service.proto:14-50) with ``LongRunningService.Get`` for its operations — wire-compatible with the
reference's stubs.  Control messages only: payload bytes move by the CUDA kernels; ``Read`` exists
for *legacy* CPU consumers and streams 2 MiB chunks + EOS out of HBM."""
from .protos import messages, method_types
from .server import ChannelManagerService, LongRunningService, PeerApiService, SlotsApiService, WorkerApiService, serve
from .pump import RemoteActionPump
from .client import ChannelManagerClient, ChannelManagerPrivateClient, LongRunningClient, PeerApiClient, SlotsApiClient, WorkerApiClient

__all__ = ["messages", "method_types", "ChannelManagerService", "SlotsApiService", "serve", "ChannelManagerClient",
           "ChannelManagerPrivateClient", "SlotsApiClient", "WorkerApiService", "LongRunningService", "WorkerApiClient",
           "LongRunningClient", "RemoteActionPump", "PeerApiService", "PeerApiClient"]
