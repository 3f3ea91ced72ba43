/*
 * lzyb_oracle.c — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product
 * (lzy_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it, as the checker.
 *
 * Plain-C, single-threaded restatement of the byte semantics of the reference's slot data plane:
 *
 *   oracle_stream_copy   — the slot→slot byte path: producer reads its backing store from `offset`
 *                          in 2 MiB heap buffers and emits chunks then EOS
 *                          (lzy/slots/src/main/java/ai/lzy/slots/OutputSlot.java:238-265);
 *                          the consumer writes every chunk at its running offset
 *                          (lzy/slots/src/main/java/ai/lzy/slots/transfers/SlotInputTransfer.java:95-98).
 *                          Net effect on bytes: dst[offset .. n) = src[offset .. n), chunked by 2 MiB.
 *   oracle_fs_copy       — the pylzy local runtime's storage client loop: 8096-byte read/write
 *                          (pylzy/lzy/storage/async_/fs.py:15-38).
 *   oracle_digest        — scalar statement of include/lzyb_digest.h (the data plane's replacement
 *                          for the streaming MD5 of pylzy/lzy/api/v1/utils/hashing.py:36-38).
 *   oracle_pattern       — the counter-based fill of BASELINE config C3.
 *   oracle_gather2d      — row gather (what serialzy does for a non-contiguous ndarray: it
 *                          serialises the logical C-order bytes).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lzyb_digest.h"

#define ORACLE_GRPC_CHUNK (2u << 20) /* OutputSlot.java:241  ByteBuffer.allocate(2 << 20L) */
#define ORACLE_FS_CHUNK 8096u        /* fs.py:19,33 */

/* returns number of chunks "sent" (excluding the EOS control message) */
uint64_t oracle_stream_copy(const uint8_t* src, uint8_t* dst, uint64_t nbytes, uint64_t offset,
                            uint64_t fail_after /* UINT64_MAX = never */, uint64_t* delivered) {
    uint8_t* heap = (uint8_t*)malloc(ORACLE_GRPC_CHUNK);
    uint64_t pos = offset, chunks = 0, read_since_open = 0;
    while (pos < nbytes) {
        if (fail_after != UINT64_MAX && read_since_open >= fail_after) break; /* FailingInputStream, SlotsTest.java:925-929 */
        uint64_t n = nbytes - pos < ORACLE_GRPC_CHUNK ? nbytes - pos : ORACLE_GRPC_CHUNK;
        memcpy(heap, src + pos, n);  /* source.read(buffer) */
        memcpy(dst + pos, heap, n);  /* outputStream.write(chunk) at currentOffset */
        pos += n;
        read_since_open += n;
        chunks++;
    }
    free(heap);
    if (delivered) *delivered = pos;
    return chunks;
}

uint64_t oracle_fs_copy(const uint8_t* src, uint8_t* dst, uint64_t nbytes) {
    uint64_t pos = 0, iters = 0;
    while (pos < nbytes) {
        uint64_t n = nbytes - pos < ORACLE_FS_CHUNK ? nbytes - pos : ORACLE_FS_CHUNK;
        memcpy(dst + pos, src + pos, n);
        pos += n;
        iters++;
    }
    return iters;
}

/* partial digest of payload bytes [0, nbytes) whose first byte sits at payload offset
 * `byte_offset` (multiple of 8): accumulates into sum / xr */
void oracle_digest_partial(const uint8_t* p, uint64_t nbytes, uint64_t byte_offset, uint64_t* sum, uint64_t* xr) {
    uint64_t i = byte_offset >> 3, s = *sum, x = *xr;
    uint64_t full = nbytes >> 3;
    for (uint64_t k = 0; k < full; ++k) {
        uint64_t w;
        memcpy(&w, p + 8 * k, 8); /* little-endian host */
        uint64_t m = lzyb_word_term(w, i + k);
        s += m;
        x ^= m;
    }
    uint64_t tail = nbytes & 7;
    if (tail) {
        uint64_t w = 0;
        memcpy(&w, p + 8 * full, tail);
        uint64_t m = lzyb_word_term(w, i + full);
        s += m;
        x ^= m;
    }
    *sum = s;
    *xr = x;
}

void oracle_digest(const uint8_t* p, uint64_t nbytes, uint64_t out[2]) {
    uint64_t s = 0, x = 0;
    oracle_digest_partial(p, nbytes, 0, &s, &x);
    lzyb_digest_finalize(s, x, nbytes, out);
}

void oracle_digest_finalize(uint64_t sum, uint64_t xr, uint64_t nbytes, uint64_t out[2]) {
    lzyb_digest_finalize(sum, xr, nbytes, out);
}

void oracle_pattern(uint8_t* dst, uint64_t nbytes, uint64_t seed) {
    for (uint64_t b = 0; b < nbytes; b += 8) {
        uint64_t w = lzyb_mix64((seed + (b >> 3) + 1ull) * LZYB_K_POS);
        uint64_t n = nbytes - b < 8 ? nbytes - b : 8;
        memcpy(dst + b, &w, n);
    }
}

void oracle_gather2d(const uint8_t* src, uint8_t* dst, uint64_t rows, uint64_t row_bytes, uint64_t pitch) {
    for (uint64_t r = 0; r < rows; ++r) memcpy(dst + r * row_bytes, src + r * pitch, row_bytes);
}
