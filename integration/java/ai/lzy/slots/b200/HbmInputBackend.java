package ai.lzy.slots.b200;

import ai.lzy.slots.backends.InputSlotBackend;
import ai.lzy.slots.backends.OutputSlotBackend;

import java.io.IOException;
import java.nio.channels.SeekableByteChannel;

/**
 * {@link InputSlotBackend} over an HBM slot frame; takes the place of {@code FileInputBackend}
 * (lzy/slots/src/main/java/ai/lzy/slots/backends/FileInputBackend.java:15-55).  The frame is opened with the
 * size the peer announced; {@code openChannel()} hands out the marker the transfers write through; after
 * delivery {@code toOutput()} turns the very same frame into the "-out" producer
 * (InputSlot.java:160-168) without copying it.
 */
public final class HbmInputBackend implements InputSlotBackend {
    private final long handle;
    private final long nbytes;
    private boolean handedOver;
    private boolean closed;

    public HbmInputBackend(String slotId, int gpu, long nbytes) throws IOException {
        this.handle = Native.check(Native.inputOpen(slotId, gpu, nbytes));
        this.nbytes = nbytes;
    }

    public long handle() {
        return handle;
    }

    @Override
    public SeekableByteChannel openChannel() {
        return new HbmSinkMarker(handle, nbytes);
    }

    @Override
    public synchronized OutputSlotBackend toOutput() {
        if (Native.inputToOutput(handle) != Native.OK) {
            throw new IllegalStateException("lzy_b200: " + Native.lastError());
        }
        handedOver = true;
        return new HbmOutputBackend(handle);
    }

    @Override
    public synchronized void close() {
        if (!closed && !handedOver) {
            Native.slotClose(handle);
        }
        closed = true;
    }
}
