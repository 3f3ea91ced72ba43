package ai.lzy.slots.b200;

import ai.lzy.slots.transfers.InputTransfer;
import ai.lzy.v1.common.LC;

import java.io.IOException;
import java.nio.channels.SeekableByteChannel;
import java.util.function.Function;

/**
 * {@link InputTransfer} for a producer whose frame lives in HBM (peer URL scheme {@code cuda+ipc://}); takes
 * the place of {@code SlotInputTransfer} (lzy/slots/src/main/java/ai/lzy/slots/transfers/SlotInputTransfer.java:21-112).
 * One {@code transferChunkTo} = one consumer-pull (CUDA kernel or copy engine + trailing digest scans) of up to
 * {@code chunkBytes}; the offset resumes where the sink stands, as the reference's does (:95-98); EOS verifies the
 * delivered bytes against the producer's digest and a mismatch is a {@link ReadException} — the consumer asks the
 * channel manager for another peer (InputSlot.java:232-260).
 */
public final class NvlinkInputTransfer implements InputTransfer {
    private final String peerId;
    private final long peer;
    private final long chunkBytes;

    /**
     * @param resolve maps the peer description to a local handle: {@code Native.slotFind(peerId)} when producer and
     *                consumer share a worker process, otherwise {@code Native.peerImport(descriptor, gpu)} with the
     *                256-byte descriptor fetched from the producer's worker ({@code Native.slotExport})
     */
    public NvlinkInputTransfer(LC.PeerDescription peer, Function<LC.PeerDescription, Long> resolve, long chunkBytes)
        throws ReadException
    {
        this.peerId = peer.getPeerId();
        long h = resolve.apply(peer);
        if (h < 0) {
            throw new ReadException("Cannot reach peer " + peerId + ": " + Native.lastError(), null);
        }
        this.peer = h;
        this.chunkBytes = chunkBytes <= 0 ? Long.MAX_VALUE : chunkBytes;
    }

    @Override
    public int transferChunkTo(SeekableByteChannel sink) throws ReadException, IOException {
        if (!(sink instanceof HbmSinkMarker marker)) {
            throw new IOException("NvlinkInputTransfer needs an HBM-backed sink");
        }
        long offset = marker.position();
        long n = Native.transferChunk(marker.handle(), peer, offset, chunkBytes);
        if (n == Native.EOF) {
            int rc = Native.inputFinish(marker.handle(), peer);
            if (rc == Native.E_PEER) {
                marker.position(0);  // the frame was reset: the next peer starts from the beginning
                throw new ReadException("Peer " + peerId + " delivered corrupt data: " + Native.lastError(), null);
            }
            if (rc != Native.OK) {
                throw new IOException("lzy_b200: " + Native.lastError());
            }
            return -1;
        }
        if (n == Native.E_PEER) {
            throw new ReadException("Cannot read from peer " + peerId + ": " + Native.lastError(), null);
        }
        if (n < 0) {
            throw new IOException("lzy_b200: " + Native.lastError());
        }
        marker.position(offset + n);
        return (int) Math.min(n, Integer.MAX_VALUE);
    }

    @Override
    public void close() {
        // an imported peer mapping is cached by the library; a same-process handle belongs to its producer
    }
}
