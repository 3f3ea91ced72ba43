/*
 * lzyb_digest.h — the slot content digest, defined once for host and device.
 *
 * The reference hashes every serialized payload with a single-threaded streaming MD5
 * (pylzy/lzy/api/v1/utils/hashing.py:36-38, used by DefaultSnapshot.put_data,
 * pylzy/lzy/api/v1/snapshot.py:147-155) and verifies nothing on the slot→slot path
 * (lzy/slots/.../transfers/SlotInputTransfer.java:95-98 writes whatever arrives).
 * MD5 is a serial chain and cannot be computed at HBM speed, so the B200 data plane
 * uses an order-independent, position-keyed 128-bit digest instead:
 *
 *   payload is read as little-endian 64-bit words w_i (i = byte_offset / 8); the last
 *   word is zero-padded.  Each word contributes
 *       m_i = mix64(w_i ^ ((i + 1) * LZYB_K_POS))
 *   and a range [a, b) of words has the partial digest
 *       sum = Σ m_i  (mod 2^64)          xr = XOR m_i
 *   Partials over disjoint ranges combine by (+, ^) — this is what lets chunked,
 *   offset-resumed transfers (InputSlot.download, lzy/slots/.../InputSlot.java:223-295)
 *   and thousands of CUDA blocks accumulate one digest with two atomics per block.
 *   The final 128-bit value folds in the byte length (lzyb_digest_finalize).
 *
 * Everything here is integer arithmetic: GPU and CPU results are bit-identical.
 */
#ifndef LZYB_DIGEST_H
#define LZYB_DIGEST_H

#include <stdint.h>

#if defined(__CUDACC__)
#define LZYB_HD __host__ __device__ __forceinline__
#else
#define LZYB_HD static inline
#endif

#define LZYB_K_POS 0x9E3779B97F4A7C15ull
#define LZYB_K_M1 0xFF51AFD7ED558CCDull
#define LZYB_K_M2 0xC4CEB9FE1A85EC53ull
#define LZYB_K_LEN 0xD6E8FEB86659FD93ull

LZYB_HD uint64_t lzyb_mix64(uint64_t x) {
    x *= LZYB_K_M1;
    x ^= x >> 32;
    x *= LZYB_K_M2;
    x ^= x >> 29;
    return x;
}

/* contribution of 64-bit word `w` sitting at word index `i` of the payload */
LZYB_HD uint64_t lzyb_word_term(uint64_t w, uint64_t i) {
    return lzyb_mix64(w ^ ((i + 1ull) * LZYB_K_POS));
}

/* fold (sum, xr, nbytes) into the published 128-bit digest {hi, lo} */
LZYB_HD void lzyb_digest_finalize(uint64_t sum, uint64_t xr, uint64_t nbytes, uint64_t out[2]) {
    uint64_t len = (nbytes + 1ull) * LZYB_K_LEN;
    uint64_t hi = lzyb_mix64(sum ^ len);
    uint64_t lo = lzyb_mix64(xr + ((sum << 32) | (sum >> 32)) + len);
    out[0] = hi;
    out[1] = lo ^ (hi >> 1);
}

/*
 * Content address of a framed VALUE.  The payload digest is blind to how the bytes are to be read: dtype,
 * serializer flags, rank and shape live in the frame header, so zeros(4, f32) and zeros(4, i32), or one buffer
 * viewed as (2, 8) and as (4, 4), share a payload digest.  The reference hashes the serialized stream, which
 * carries dtype and shape (pylzy/lzy/api/v1/snapshot.py:147-155), so its md5 tells such values apart; the key
 * below does the same for a frame: an order-dependent chain over (dtype|flags, ndim, shape[0..ndim)) folded into
 * both halves of the digest.  Storage URIs, `data_hash` and hence `cache=True` result URIs are built from this
 * key, never from the bare payload digest.
 *
 * Scope: like the digest itself this is an error-detecting / de-duplicating code over honest data, not a
 * cryptographic hash — (sum, xor) of bijectively mixed words admits constructed second preimages.
 */
#define LZYB_K_META 0xA0761D6478BD642Full

LZYB_HD void lzyb_content_key_fold(const uint64_t digest[2], uint32_t dtype_flags, uint32_t ndim, const uint64_t* shape,
                              uint64_t out[2]) {
    uint64_t h = lzyb_mix64((((uint64_t)dtype_flags << 32) | (uint64_t)ndim) ^ LZYB_K_META);
    for (uint32_t i = 0; i < ndim; ++i) h = lzyb_mix64(h ^ lzyb_mix64(shape[i] + ((uint64_t)i + 1ull) * LZYB_K_POS));
    uint64_t hi = lzyb_mix64(digest[0] ^ h);
    uint64_t lo = lzyb_mix64(digest[1] + ((h << 32) | (h >> 32)) + LZYB_K_LEN);
    out[0] = hi;
    out[1] = lo ^ (hi >> 1);
}

#endif /* LZYB_DIGEST_H */
