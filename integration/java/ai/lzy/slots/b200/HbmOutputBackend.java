package ai.lzy.slots.b200;

import ai.lzy.slots.backends.OutputSlotBackend;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.channels.ReadableByteChannel;

/**
 * {@link OutputSlotBackend} over an HBM slot frame; takes the place of {@code OutputPipeBackend}
 * (lzy/slots/src/main/java/ai/lzy/slots/backends/OutputPipeBackend.java:18-69).  The op writes its result into
 * the frame ({@link #devicePtr()}) or hands a device pointer to {@link #pack}; {@link #waitCompleted()} is the
 * commit: one CUDA launch frames the payload and publishes its 128-bit digest.
 */
public final class HbmOutputBackend implements OutputSlotBackend {
    private final long handle;
    private final long[] digest = new long[2];
    private boolean closed;

    public HbmOutputBackend(String slotId, int gpu, long nbytes) throws IOException {
        handle = Native.check(Native.slotCreate(slotId, gpu, nbytes));
    }

    /** an input slot that was handed over after delivery ({@code InputSlotBackend.toOutput()}) */
    HbmOutputBackend(long committedHandle) {
        handle = committedHandle;
    }

    public long handle() {
        return handle;
    }

    public long devicePtr() throws IOException {
        return Native.check(Native.slotDevicePtr(handle));
    }

    public void pack(long sourceDevicePtr, long nbytes) throws IOException {
        Native.check(Native.slotPack(handle, sourceDevicePtr, nbytes));
    }

    public void packHost(ByteBuffer direct, long nbytes) throws IOException {
        Native.check(Native.slotPackHost(handle, direct, nbytes));
    }

    @Override
    public void waitCompleted() throws IOException {
        Native.check(Native.slotCommit(handle, digest));
    }

    /** legacy consumers ({@code LzySlotsApi.Read}): bytes leave HBM through a direct bounce buffer */
    @Override
    public ReadableByteChannel readFromOffset(long offset) {
        return new ReadableByteChannel() {
            private long pos = offset;
            private boolean open = true;

            @Override
            public int read(ByteBuffer dst) throws IOException {
                if (!dst.hasRemaining()) {
                    return 0;
                }
                ByteBuffer direct = dst.isDirect() ? dst.slice() : ByteBuffer.allocateDirect(dst.remaining());
                long n = Native.check(Native.slotRead(handle, pos, direct, direct.capacity()));
                if (n == Native.EOF) {
                    return -1;
                }
                if (dst.isDirect()) {
                    dst.position(dst.position() + (int) n);
                } else {
                    direct.limit((int) n);
                    dst.put(direct);
                }
                pos += n;
                return (int) n;
            }

            @Override
            public boolean isOpen() {
                return open;
            }

            @Override
            public void close() {
                open = false;
            }
        };
    }

    @Override
    public synchronized void close() {
        if (!closed) {
            closed = true;
            Native.slotClose(handle);
        }
    }
}
