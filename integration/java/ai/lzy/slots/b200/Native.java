package ai.lzy.slots.b200;

import java.nio.ByteBuffer;

/**
 * The C ABI of liblzy_b200.so (include/lzy_b200.h) as Java sees it, through integration/jni/lzy_b200_jni.cpp.
 * Handles are opaque longs; a negative return value is an LZYB_E_* code (0 = OK, -1 = EOF) and
 * {@link #lastError()} holds the message of the last failure on the calling thread.
 */
public final class Native {
    static {
        System.loadLibrary("lzy_b200_jni");
    }

    private Native() {}

    public static final long OK = 0;
    public static final long EOF = -1;
    /** the peer could not deliver (a ReadException: ask the channel manager for another producer) */
    public static final long E_PEER = -2;

    public static native int init(int firstGpu, int nGpus, long hostRingBytes);

    public static native void shutdown();

    public static native String lastError();

    // OutputSlotBackend
    public static native long slotCreate(String slotId, int gpu, long nbytes);

    public static native long slotDevicePtr(long handle);

    public static native int slotPack(long handle, long devicePtr, long nbytes);

    public static native int slotPackHost(long handle, ByteBuffer direct, long nbytes);

    public static native int slotCommit(long handle, long[] digestOut);

    public static native long slotRead(long handle, long offset, ByteBuffer direct, long len);

    public static native int slotClose(long handle);

    public static native long slotFind(String slotId);

    public static native long slotBytes(long handle);

    // InputSlotBackend + InputTransfer
    public static native long inputOpen(String slotId, int gpu, long nbytes);

    public static native long transferChunk(long dst, long srcPeer, long offset, long maxLen);

    public static native int inputFinish(long dst, long srcPeer);

    public static native long inputWriteHost(long dst, long offset, ByteBuffer direct, long len);

    public static native int inputFinishExpect(long dst, byte[] digest16);

    public static native int inputToOutput(long handle);

    public static native int inputReset(long handle);

    // peers in other worker processes
    public static native byte[] slotExport(long handle);

    public static native long peerImport(byte[] descriptor, int ontoGpu);

    // producer-issued push between worker processes (include/lzy_b200.h: lzyb_sink_export / lzyb_slot_push /
    // lzyb_input_expect_push): the consumer describes its empty input frame, the producer's process pushes into it
    public static native byte[] sinkExport(long inputHandle);

    public static native int slotPush(long srcHandle, byte[] sinkDescriptor);

    public static native long inputExpectPush(long dst, byte[] producerDescriptor);

    // the channel has exactly one consumer and no other reader: a same-GPU consumer takes the frame over (toOutput-style)
    public static native int slotSetMovable(long handle, boolean movable);

    static long check(long rc) throws java.io.IOException {
        if (rc < 0 && rc != EOF) {
            throw new java.io.IOException("lzy_b200: " + lastError() + " (code " + rc + ")");
        }
        return rc;
    }
}
