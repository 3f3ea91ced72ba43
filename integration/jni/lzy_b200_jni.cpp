// lzy_b200_jni.cpp — the JNI shim between the reference's Java worker (lzy/slots) and liblzy_b200.so.
//
// One function per `native` method of ai.lzy.slots.b200.Native (integration/java/...); no logic beyond argument
// marshalling: a Java String becomes a C string, a direct ByteBuffer its address, a handle a jlong.  Errors travel
// as negative return values (include/lzy_b200.h: LZYB_E_*) and the message through lastError(), so the Java side
// decides between ReadException (LZYB_E_PEER: ask the channel manager for another peer) and IOException
// (everything else: the slot fails) exactly where the reference's transfers do
// (lzy/slots/src/main/java/ai/lzy/slots/transfers/SlotInputTransfer.java:43-99).
//
// Build on a box with a JDK (this image has none; tests/test_integration_sources.py type-checks the file against
// a stand-in jni.h):
//   g++ -std=c++17 -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude
//       integration/jni/lzy_b200_jni.cpp -Llzy_b200/lib -llzy_b200 -o liblzy_b200_jni.so
#include <jni.h>

#include <cstring>

#include "lzy_b200.h"

namespace {

// RAII view of a Java String as UTF-8
class Utf {
public:
    Utf(JNIEnv* env, jstring s) : env_(env), s_(s), p_(s ? env->GetStringUTFChars(s, nullptr) : nullptr) {}
    ~Utf() {
        if (p_) env_->ReleaseStringUTFChars(s_, p_);
    }
    const char* get() const { return p_ ? p_ : ""; }

private:
    JNIEnv* env_;
    jstring s_;
    const char* p_;
};

inline jlong handle_or_error(int rc, lzyb_handle_t h) { return rc == LZYB_OK ? static_cast<jlong>(h) : static_cast<jlong>(rc); }

}  // namespace

extern "C" {

// ---- lifecycle -------------------------------------------------------------------------------------------------
JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_init(JNIEnv*, jclass, jint firstGpu, jint nGpus, jlong hostRingBytes) {
    return lzyb_init(firstGpu, nGpus, static_cast<uint64_t>(hostRingBytes));
}

JNIEXPORT void JNICALL Java_ai_lzy_slots_b200_Native_shutdown(JNIEnv*, jclass) { lzyb_shutdown(); }

JNIEXPORT jstring JNICALL Java_ai_lzy_slots_b200_Native_lastError(JNIEnv* env, jclass) {
    return env->NewStringUTF(lzyb_last_error());
}

// ---- OutputSlotBackend (backends/OutputSlotBackend.java:6-16) ------------------------------------------------------
JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_slotCreate(JNIEnv* env, jclass, jstring slotId, jint gpu, jlong nbytes) {
    Utf id(env, slotId);
    lzyb_handle_t h = 0;
    return handle_or_error(lzyb_slot_create(id.get(), gpu, static_cast<uint64_t>(nbytes), &h), h);
}

JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_slotDevicePtr(JNIEnv*, jclass, jlong handle) {
    void* p = nullptr;
    int rc = lzyb_slot_device_ptr(static_cast<lzyb_handle_t>(handle), &p);
    return rc == LZYB_OK ? reinterpret_cast<jlong>(p) : static_cast<jlong>(rc);
}

// payload already on the device at `devicePtr` (the op's result): pack + frame + digest in one launch
JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_slotPack(JNIEnv*, jclass, jlong handle, jlong devicePtr, jlong nbytes) {
    return lzyb_slot_pack(static_cast<lzyb_handle_t>(handle), reinterpret_cast<const void*>(devicePtr),
                          static_cast<uint64_t>(nbytes), nullptr);
}

// payload in a direct ByteBuffer (what the op wrote to its output pipe)
JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_slotPackHost(JNIEnv* env, jclass, jlong handle, jobject direct, jlong nbytes) {
    void* src = env->GetDirectBufferAddress(direct);
    if (!src || env->GetDirectBufferCapacity(direct) < nbytes) {
        lzyb_set_last_error("slotPackHost needs a direct ByteBuffer of at least nbytes");
        return LZYB_E_ARG;
    }
    return lzyb_slot_pack_host(static_cast<lzyb_handle_t>(handle), src, static_cast<uint64_t>(nbytes), nullptr);
}

// waitCompleted(): the digest (2 longs) lands in `digestOut`
JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_slotCommit(JNIEnv* env, jclass, jlong handle, jlongArray digestOut) {
    uint64_t digest[2] = {0, 0};
    int rc = lzyb_slot_commit(static_cast<lzyb_handle_t>(handle), digest, nullptr);
    if (rc == LZYB_OK && digestOut && env->GetArrayLength(digestOut) >= 2) {
        const jlong out[2] = {static_cast<jlong>(digest[0]), static_cast<jlong>(digest[1])};
        env->SetLongArrayRegion(digestOut, 0, 2, out);
    }
    return rc;
}

// readFromOffset(): bytes [offset, offset + len) into a direct ByteBuffer; returns bytes read, LZYB_EOF or an error
JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_slotRead(JNIEnv* env, jclass, jlong handle, jlong offset, jobject direct, jlong len) {
    void* dst = env->GetDirectBufferAddress(direct);
    if (!dst || env->GetDirectBufferCapacity(direct) < len) {
        lzyb_set_last_error("slotRead needs a direct ByteBuffer of at least len bytes");
        return LZYB_E_ARG;
    }
    return static_cast<jlong>(lzyb_slot_read(static_cast<lzyb_handle_t>(handle), static_cast<uint64_t>(offset), dst,
                                             static_cast<uint64_t>(len)));
}

JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_slotClose(JNIEnv*, jclass, jlong handle) {
    return lzyb_slot_close(static_cast<lzyb_handle_t>(handle));
}

JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_slotFind(JNIEnv* env, jclass, jstring slotId) {
    Utf id(env, slotId);
    lzyb_handle_t h = 0;
    return handle_or_error(lzyb_slot_find(id.get(), &h), h);
}

JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_slotBytes(JNIEnv*, jclass, jlong handle) {
    lzyb_slot_info_t info;
    int rc = lzyb_slot_info(static_cast<lzyb_handle_t>(handle), &info);
    return rc == LZYB_OK ? static_cast<jlong>(info.nbytes) : static_cast<jlong>(rc);
}

// ---- InputSlotBackend (backends/InputSlotBackend.java:6-13) + InputTransfer (transfers/InputTransfer.java:6-20) -------
JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_inputOpen(JNIEnv* env, jclass, jstring slotId, jint gpu, jlong nbytes) {
    Utf id(env, slotId);
    lzyb_handle_t h = 0;
    return handle_or_error(lzyb_input_open(id.get(), gpu, static_cast<uint64_t>(nbytes), &h), h);
}

// transferChunkTo(): one pull of at most maxLen bytes at `offset`; bytes moved, LZYB_EOF, LZYB_E_PEER or another error
JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_transferChunk(JNIEnv*, jclass, jlong dst, jlong srcPeer, jlong offset, jlong maxLen) {
    return static_cast<jlong>(lzyb_transfer_chunk(static_cast<lzyb_handle_t>(dst), static_cast<lzyb_handle_t>(srcPeer),
                                                  static_cast<uint64_t>(offset), static_cast<uint64_t>(maxLen), nullptr));
}

// EOS: verify what landed against the producer's digest
JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_inputFinish(JNIEnv*, jclass, jlong dst, jlong srcPeer) {
    return lzyb_input_finish(static_cast<lzyb_handle_t>(dst), static_cast<lzyb_handle_t>(srcPeer), nullptr);
}

// bytes from a storage peer (StorageInputTransfer): host buffer → frame, digest accumulated on the way
JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_inputWriteHost(JNIEnv* env, jclass, jlong dst, jlong offset, jobject direct, jlong len) {
    void* src = env->GetDirectBufferAddress(direct);
    if (!src || env->GetDirectBufferCapacity(direct) < len) {
        lzyb_set_last_error("inputWriteHost needs a direct ByteBuffer of at least len bytes");
        return LZYB_E_ARG;
    }
    return static_cast<jlong>(lzyb_input_write_host(static_cast<lzyb_handle_t>(dst), static_cast<uint64_t>(offset), src,
                                                    static_cast<uint64_t>(len), nullptr));
}

// the storage object carries no digest of its own (digest == null) or the uploader's (2 longs)
JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_inputFinishExpect(JNIEnv* env, jclass, jlong dst, jbyteArray digest16) {
    if (!digest16) return lzyb_input_finish_expect(static_cast<lzyb_handle_t>(dst), nullptr, nullptr);
    if (env->GetArrayLength(digest16) != 16) {
        lzyb_set_last_error("digest must be 16 bytes");
        return LZYB_E_ARG;
    }
    uint64_t d[2];
    env->GetByteArrayRegion(digest16, 0, 16, reinterpret_cast<jbyte*>(d));
    return lzyb_input_finish_expect(static_cast<lzyb_handle_t>(dst), d, nullptr);
}

JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_inputToOutput(JNIEnv*, jclass, jlong handle) {
    return lzyb_input_to_output(static_cast<lzyb_handle_t>(handle));
}

JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_inputReset(JNIEnv*, jclass, jlong handle) {
    return lzyb_input_reset(static_cast<lzyb_handle_t>(handle));
}

// ---- peers in other worker processes: CUDA IPC ---------------------------------------------------------------------
// the 256-byte descriptor a producer publishes behind its "cuda+ipc://" peer URL
JNIEXPORT jbyteArray JNICALL Java_ai_lzy_slots_b200_Native_slotExport(JNIEnv* env, jclass, jlong handle) {
    lzyb_ipc_desc_t desc;
    if (lzyb_slot_export(static_cast<lzyb_handle_t>(handle), &desc) != LZYB_OK) return nullptr;
    jbyteArray out = env->NewByteArray(static_cast<jsize>(sizeof(desc)));
    if (out) env->SetByteArrayRegion(out, 0, static_cast<jsize>(sizeof(desc)), reinterpret_cast<const jbyte*>(&desc));
    return out;
}

JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_peerImport(JNIEnv* env, jclass, jbyteArray descBytes, jint ontoGpu) {
    lzyb_ipc_desc_t desc;
    if (!descBytes || env->GetArrayLength(descBytes) != static_cast<jsize>(sizeof(desc))) {
        lzyb_set_last_error("peer descriptor must be 256 bytes");
        return LZYB_E_ARG;
    }
    env->GetByteArrayRegion(descBytes, 0, static_cast<jsize>(sizeof(desc)), reinterpret_cast<jbyte*>(&desc));
    lzyb_handle_t h = 0;
    return handle_or_error(lzyb_peer_import(&desc, ontoGpu, &h), h);
}

// ---- producer-issued push between worker processes ------------------------------------------------------------------
JNIEXPORT jbyteArray JNICALL Java_ai_lzy_slots_b200_Native_sinkExport(JNIEnv* env, jclass, jlong inputHandle) {
    lzyb_ipc_desc_t desc;
    if (lzyb_sink_export(static_cast<lzyb_handle_t>(inputHandle), &desc) != LZYB_OK) return nullptr;
    jbyteArray out = env->NewByteArray(static_cast<jsize>(sizeof(desc)));
    if (out) env->SetByteArrayRegion(out, 0, static_cast<jsize>(sizeof(desc)), reinterpret_cast<const jbyte*>(&desc));
    return out;
}

JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_slotPush(JNIEnv* env, jclass, jlong srcHandle, jbyteArray sinkBytes) {
    lzyb_ipc_desc_t sink;
    if (!sinkBytes || env->GetArrayLength(sinkBytes) != static_cast<jsize>(sizeof(sink))) {
        lzyb_set_last_error("sink descriptor must be 256 bytes");
        return LZYB_E_ARG;
    }
    env->GetByteArrayRegion(sinkBytes, 0, static_cast<jsize>(sizeof(sink)), reinterpret_cast<jbyte*>(&sink));
    return lzyb_slot_push(static_cast<lzyb_handle_t>(srcHandle), &sink, 0, 0, nullptr);
}

JNIEXPORT jlong JNICALL Java_ai_lzy_slots_b200_Native_inputExpectPush(JNIEnv* env, jclass, jlong dst, jbyteArray producerBytes) {
    lzyb_ipc_desc_t producer;
    if (!producerBytes || env->GetArrayLength(producerBytes) != static_cast<jsize>(sizeof(producer))) {
        lzyb_set_last_error("producer descriptor must be 256 bytes");
        return LZYB_E_ARG;
    }
    env->GetByteArrayRegion(producerBytes, 0, static_cast<jsize>(sizeof(producer)), reinterpret_cast<jbyte*>(&producer));
    return static_cast<jlong>(lzyb_input_expect_push(static_cast<lzyb_handle_t>(dst), &producer, 0, 0, nullptr));
}

JNIEXPORT jint JNICALL Java_ai_lzy_slots_b200_Native_slotSetMovable(JNIEnv*, jclass, jlong handle, jboolean movable) {
    return lzyb_slot_set_movable(static_cast<lzyb_handle_t>(handle), movable ? 1 : 0);
}

}  // extern "C"
