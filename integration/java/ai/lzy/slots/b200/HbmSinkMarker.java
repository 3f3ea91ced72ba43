package ai.lzy.slots.b200;

import java.nio.ByteBuffer;
import java.nio.channels.SeekableByteChannel;

/**
 * What {@code InputSlotBackend.openChannel()} returns for an HBM-backed input slot: the reference's InputSlot
 * only hands this object to {@code InputTransfer.transferChunkTo(sink)} and positions it
 * (lzy/slots/src/main/java/ai/lzy/slots/InputSlot.java:223-295), so it carries the frame's handle and the
 * running offset; payload bytes never enter the JVM.  A transfer that does not know about HBM frames (a stock
 * {@code SlotInputTransfer} reading a CPU peer over gRPC) can still {@code write()} into it: the bytes are staged
 * through a direct buffer into the frame.
 */
public final class HbmSinkMarker implements SeekableByteChannel {
    private final long handle;
    private final long size;
    private long position;
    private boolean open = true;

    HbmSinkMarker(long handle, long size) {
        this.handle = handle;
        this.size = size;
    }

    public long handle() {
        return handle;
    }

    @Override
    public int write(ByteBuffer src) throws java.io.IOException {
        int n = src.remaining();
        ByteBuffer direct = src.isDirect() ? src.slice() : ByteBuffer.allocateDirect(n).put(src.duplicate()).flip();
        long moved = Native.check(Native.inputWriteHost(handle, position, direct, n));
        src.position(src.position() + (int) moved);
        position += moved;
        return (int) moved;
    }

    @Override
    public int read(ByteBuffer dst) {
        throw new java.nio.channels.NonReadableChannelException();
    }

    @Override
    public long position() {
        return position;
    }

    @Override
    public SeekableByteChannel position(long newPosition) {
        position = newPosition;
        return this;
    }

    @Override
    public long size() {
        return size;
    }

    @Override
    public SeekableByteChannel truncate(long newSize) throws java.io.IOException {
        if (newSize == 0) {
            Native.check(Native.inputReset(handle));
            position = 0;
        }
        return this;
    }

    @Override
    public boolean isOpen() {
        return open;
    }

    @Override
    public void close() {
        open = false;
    }
}
