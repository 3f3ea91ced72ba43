/*
 * hostmem_backend.c — TEST INFRASTRUCTURE (never part of the product library).
 *
 * The reference tests its slot state machines with an in-memory backend that has failure knobs
 * (SlotsTest.InMemBackend: failOpen / failRead / failWaitCompleted / failAfter,
 * lzy/slots/src/test/java/ai/lzy/slots/SlotsTest.java:866-943).  This is the same thing for the C++
 * slots runtime: a lzyb_backend_ops_t over malloc'ed memory, installed with lzyb_slots_set_backend()
 * by tests/ so the InputSlot / OutputSlot / execution-context logic runs on a box without a GPU.
 * Digests use the shared definition (include/lzyb_digest.h) so verification behaves like the HBM path.
 */
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/lzy_b200_host.h"
#include "../../include/lzyb_digest.h"

#define MAX_SLOTS 4096

typedef struct {
    int used;
    lzyb_handle_t h;
    char id[256];
    uint8_t* data;
    uint64_t nbytes, delivered;
    int state, is_input, gpu;
    uint64_t digest[2];
    uint32_t dtype, ndim;
    uint64_t shape[LZYB_MAX_DIMS];
    long long fail_after; /* -1 = never */
    int fail_read;
    int refs;
    int mapped;          /* data is an mmap of another process's export (peer_import) */
    char shm_name[64];   /* this slot has been exported under that POSIX shm name */
} mem_slot_t;

static mem_slot_t g_slots[MAX_SLOTS];
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static lzyb_handle_t g_next = 0x7000000000000001ull;
static int g_fail_open = 0, g_fail_commit = 0;

static int err(int code, const char* msg) {
    lzyb_set_last_error(msg);
    return code;
}

static mem_slot_t* find(lzyb_handle_t h) {
    for (int i = 0; i < MAX_SLOTS; ++i)
        if (g_slots[i].used && g_slots[i].h == h) return &g_slots[i];
    return NULL;
}

static void digest_of(const uint8_t* p, uint64_t n, uint64_t out[2]) {
    uint64_t s = 0, x = 0, full = n >> 3;
    for (uint64_t k = 0; k < full; ++k) {
        uint64_t w;
        memcpy(&w, p + 8 * k, 8);
        uint64_t m = lzyb_word_term(w, k);
        s += m;
        x ^= m;
    }
    if (n & 7) {
        uint64_t w = 0;
        memcpy(&w, p + 8 * full, n & 7);
        uint64_t m = lzyb_word_term(w, full);
        s += m;
        x ^= m;
    }
    lzyb_digest_finalize(s, x, n, out);
}

static int new_slot(const char* id, int gpu, uint64_t nbytes, int is_input, lzyb_handle_t* out) {
    pthread_mutex_lock(&g_mu);
    if (is_input && g_fail_open) {
        g_fail_open = 0;
        pthread_mutex_unlock(&g_mu);
        return err(LZYB_E_LOCAL, "Cannot open channel");
    }
    for (int i = 0; i < MAX_SLOTS; ++i) {
        if (!g_slots[i].used) {
            mem_slot_t* s = &g_slots[i];
            memset(s, 0, sizeof(*s));
            s->used = 1;
            s->h = g_next++;
            snprintf(s->id, sizeof(s->id), "%s", id);
            s->data = (uint8_t*)calloc(nbytes ? nbytes : 1, 1);
            s->nbytes = nbytes;
            s->is_input = is_input;
            s->gpu = gpu;
            s->fail_after = -1;
            s->refs = 1;
            *out = s->h;
            pthread_mutex_unlock(&g_mu);
            return LZYB_OK;
        }
    }
    pthread_mutex_unlock(&g_mu);
    return err(LZYB_E_LOCAL, "too many slots");
}

static int m_slot_create(const char* id, int gpu, uint64_t n, lzyb_handle_t* out) { return new_slot(id, gpu, n, 0, out); }
static int m_input_open(const char* id, int gpu, uint64_t n, lzyb_handle_t* out) { return new_slot(id, gpu, n, 1, out); }

static int m_slot_commit(lzyb_handle_t h, uint64_t dg[2], void* stream) {
    (void)stream;
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    if (!s) {
        pthread_mutex_unlock(&g_mu);
        return err(LZYB_E_ARG, "unknown slot handle");
    }
    if (g_fail_commit) {
        g_fail_commit = 0;
        pthread_mutex_unlock(&g_mu);
        return err(LZYB_E_LOCAL, "Cannot wait");
    }
    if (s->state != LZYB_ST_READY) {
        digest_of(s->data, s->nbytes, s->digest);
        s->state = LZYB_ST_READY;
        s->delivered = s->nbytes;
    }
    if (dg) {
        dg[0] = s->digest[0];
        dg[1] = s->digest[1];
    }
    pthread_mutex_unlock(&g_mu);
    return LZYB_OK;
}

static int64_t m_slot_read(lzyb_handle_t h, uint64_t off, void* dst, uint64_t len) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    int64_t rc;
    if (!s)
        rc = err(LZYB_E_ARG, "unknown slot handle");
    else if (s->state != LZYB_ST_READY)
        rc = err(LZYB_E_STATE, "slot is not ready");
    else if (s->is_input)
        rc = err(LZYB_E_UNIMPLEMENTED, "unexpected `read` call");
    else if (s->fail_read)
        rc = err(LZYB_E_LOCAL, "Cannot read");
    else if (off >= s->nbytes)
        rc = LZYB_EOF;
    else {
        uint64_t n = s->nbytes - off < len ? s->nbytes - off : len;
        memcpy(dst, s->data + off, n);
        rc = (int64_t)n;
    }
    pthread_mutex_unlock(&g_mu);
    return rc;
}

static int m_slot_info(lzyb_handle_t h, lzyb_slot_info_t* out) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    if (!s) {
        pthread_mutex_unlock(&g_mu);
        return err(LZYB_E_ARG, "unknown slot handle");
    }
    memset(out, 0, sizeof(*out));
    out->nbytes = s->nbytes;
    out->gpu = s->gpu;
    out->state = s->state;
    out->is_input = s->is_input;
    out->delivered = s->delivered;
    out->digest[0] = s->digest[0];
    out->digest[1] = s->digest[1];
    out->dtype = s->dtype;
    out->ndim = s->ndim;
    memcpy(out->shape, s->shape, sizeof(out->shape));
    pthread_mutex_unlock(&g_mu);
    return LZYB_OK;
}

static int m_slot_set_meta(lzyb_handle_t h, uint32_t dtype, uint32_t ndim, const uint64_t* shape) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    if (s) {
        s->dtype = dtype;
        s->ndim = ndim;
        memset(s->shape, 0, sizeof(s->shape));
        for (uint32_t i = 0; i < ndim && i < LZYB_MAX_DIMS; ++i) s->shape[i] = shape[i];
    }
    pthread_mutex_unlock(&g_mu);
    return s ? LZYB_OK : err(LZYB_E_ARG, "unknown slot handle");
}

static int m_slot_find(const char* id, lzyb_handle_t* out) {
    pthread_mutex_lock(&g_mu);
    lzyb_handle_t best = 0;
    for (int i = 0; i < MAX_SLOTS; ++i)
        if (g_slots[i].used && strcmp(g_slots[i].id, id) == 0 && g_slots[i].h > best) best = g_slots[i].h;
    pthread_mutex_unlock(&g_mu);
    if (!best) return err(LZYB_E_NOT_FOUND, "Slot not found");
    *out = best;
    return LZYB_OK;
}

static int m_slot_close(lzyb_handle_t h) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    if (s && --s->refs <= 0) {
        if (s->mapped)
            munmap(s->data, s->nbytes ? s->nbytes : 1);
        else
            free(s->data);
        if (s->shm_name[0]) shm_unlink(s->shm_name);
        s->used = 0;
    }
    pthread_mutex_unlock(&g_mu);
    return LZYB_OK;
}

static int m_slot_retain(lzyb_handle_t h) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    if (s) s->refs++;
    pthread_mutex_unlock(&g_mu);
    return s ? LZYB_OK : err(LZYB_E_ARG, "unknown slot handle");
}

static long long g_moving_calls = 0; /* transfer_chunk calls that moved bytes (tests count slices with it) */
static int64_t m_transfer_chunk(lzyb_handle_t dh, lzyb_handle_t sh, uint64_t off, uint64_t max_len, void* stream) {
    (void)stream;
    pthread_mutex_lock(&g_mu);
    mem_slot_t *d = find(dh), *s = find(sh);
    int64_t rc;
    if (!d || !s)
        rc = err(LZYB_E_ARG, "unknown slot handle");
    else if (s->state != LZYB_ST_READY || s->nbytes != d->nbytes)
        rc = err(LZYB_E_PEER, "peer has no complete data");
    else if (off != d->delivered)
        rc = err(LZYB_E_ARG, "offset != delivered");
    else if (off == d->nbytes)
        rc = LZYB_EOF;
    else if (s->fail_read)
        rc = err(LZYB_E_PEER, "Cannot read");
    else {
        uint64_t len = d->nbytes - off;
        if (max_len < len) len = max_len & ~15ull;
        if (s->fail_after >= 0) { /* FailingInputStream */
            if ((long long)off >= s->fail_after) {
                pthread_mutex_unlock(&g_mu);
                return err(LZYB_E_PEER, "Cannot read");
            }
            if ((long long)(off + len) > s->fail_after) len = (uint64_t)(s->fail_after - (long long)off);
        }
        memcpy(d->data + off, s->data + off, len);
        g_moving_calls++;
        d->delivered += len;
        d->state = LZYB_ST_FILLING;
        rc = (int64_t)len;
    }
    pthread_mutex_unlock(&g_mu);
    return rc;
}

static int finish_with(mem_slot_t* d, const uint64_t* expect) {
    if (d->state == LZYB_ST_READY) return LZYB_OK;
    if (d->delivered != d->nbytes) return err(LZYB_E_STATE, "not fully delivered");
    uint64_t dg[2];
    digest_of(d->data, d->nbytes, dg);
    if (expect && (dg[0] != expect[0] || dg[1] != expect[1])) {
        d->delivered = 0;
        d->state = LZYB_ST_EMPTY;
        return err(LZYB_E_PEER, "digest mismatch");
    }
    d->digest[0] = dg[0];
    d->digest[1] = dg[1];
    d->state = LZYB_ST_READY;
    return LZYB_OK;
}

static int m_input_finish(lzyb_handle_t dh, lzyb_handle_t sh, void* stream) {
    (void)stream;
    pthread_mutex_lock(&g_mu);
    mem_slot_t *d = find(dh), *s = find(sh);
    int rc = (!d || !s) ? err(LZYB_E_ARG, "unknown slot handle") : finish_with(d, s->digest);
    pthread_mutex_unlock(&g_mu);
    return rc;
}

static int64_t m_input_write_host(lzyb_handle_t dh, uint64_t off, const void* src, uint64_t len, void* stream) {
    (void)stream;
    pthread_mutex_lock(&g_mu);
    mem_slot_t* d = find(dh);
    int64_t rc;
    if (!d)
        rc = err(LZYB_E_ARG, "unknown slot handle");
    else if (off != d->delivered)
        rc = err(LZYB_E_ARG, "offset != delivered");
    else if (off == d->nbytes)
        rc = LZYB_EOF;
    else {
        if (off + len > d->nbytes) len = d->nbytes - off;
        memcpy(d->data + off, src, len);
        d->delivered += len;
        d->state = LZYB_ST_FILLING;
        rc = (int64_t)len;
    }
    pthread_mutex_unlock(&g_mu);
    return rc;
}

static int m_input_finish_expect(lzyb_handle_t dh, const uint64_t dg[2], void* stream) {
    (void)stream;
    pthread_mutex_lock(&g_mu);
    mem_slot_t* d = find(dh);
    int rc = !d ? err(LZYB_E_ARG, "unknown slot handle") : finish_with(d, dg);
    pthread_mutex_unlock(&g_mu);
    return rc;
}

static int m_input_to_output(lzyb_handle_t h) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    int rc = LZYB_OK;
    if (!s)
        rc = err(LZYB_E_ARG, "unknown slot handle");
    else if (s->state != LZYB_ST_READY)
        rc = err(LZYB_E_STATE, "not ready");
    else
        s->is_input = 0;
    pthread_mutex_unlock(&g_mu);
    return rc;
}

static int m_input_reset(lzyb_handle_t h) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    if (s) {
        s->delivered = 0;
        s->state = LZYB_ST_EMPTY;
    }
    pthread_mutex_unlock(&g_mu);
    return s ? LZYB_OK : err(LZYB_E_ARG, "unknown slot handle");
}

/* The cross-process path of this backend, standing in for CUDA IPC: an export publishes a copy of the committed
 * payload as a POSIX shared-memory object whose name rides in the descriptor's 64 handle bytes; an import maps it
 * read-only and registers it as a READY source slot. */
static int m_slot_export(lzyb_handle_t h, lzyb_ipc_desc_t* out) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    int rc = LZYB_OK;
    if (!s || s->state != LZYB_ST_READY) {
        rc = err(LZYB_E_STATE, "slot is not committed");
    } else {
        if (!s->shm_name[0]) {
            snprintf(s->shm_name, sizeof(s->shm_name), "/lzyb-hostmem-%d-%llx", (int)getpid(), (unsigned long long)s->h);
            int fd = shm_open(s->shm_name, O_CREAT | O_RDWR | O_TRUNC, 0600);
            if (fd < 0 || ftruncate(fd, (off_t)(s->nbytes ? s->nbytes : 1)) != 0) {
                rc = err(LZYB_E_LOCAL, "cannot create shared memory");
                s->shm_name[0] = 0;
            } else if (s->nbytes) {
                void* p = mmap(NULL, s->nbytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
                if (p == MAP_FAILED) {
                    rc = err(LZYB_E_LOCAL, "cannot map shared memory");
                } else {
                    memcpy(p, s->data, s->nbytes);
                    munmap(p, s->nbytes);
                }
            }
            if (fd >= 0) close(fd);
        }
        if (rc == LZYB_OK) {
            memset(out, 0, sizeof(*out));
            memcpy(out->ipc_handle, s->shm_name, strlen(s->shm_name) + 1);
            out->buffer_bytes = s->nbytes;
            out->nbytes = s->nbytes;
            out->gpu = s->gpu;
            out->pid = (int32_t)getpid();
            out->digest[0] = s->digest[0];
            out->digest[1] = s->digest[1];
            out->dtype = s->dtype;
            out->ndim = s->ndim;
            memcpy(out->shape, s->shape, sizeof(out->shape));
        }
    }
    pthread_mutex_unlock(&g_mu);
    return rc;
}

static int m_peer_import(const lzyb_ipc_desc_t* desc, int gpu, lzyb_handle_t* out) {
    char name[65];
    memcpy(name, desc->ipc_handle, 64);
    name[64] = 0;
    int fd = shm_open(name, O_RDONLY, 0);
    if (fd < 0) return err(LZYB_E_PEER, "cannot open the peer's shared memory");
    void* p = mmap(NULL, desc->nbytes ? desc->nbytes : 1, PROT_READ, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return err(LZYB_E_PEER, "cannot map the peer's shared memory");
    pthread_mutex_lock(&g_mu);
    int rc = LZYB_E_LOCAL;
    for (int i = 0; i < MAX_SLOTS; ++i) {
        if (g_slots[i].used) continue;
        mem_slot_t* s = &g_slots[i];
        memset(s, 0, sizeof(*s));
        s->used = 1;
        s->h = g_next++;
        snprintf(s->id, sizeof(s->id), "import:%d:%s", desc->pid, name);
        s->data = (uint8_t*)p;
        s->mapped = 1;
        s->nbytes = s->delivered = desc->nbytes;
        s->state = LZYB_ST_READY;
        s->gpu = gpu;
        s->digest[0] = desc->digest[0];
        s->digest[1] = desc->digest[1];
        s->dtype = desc->dtype;
        s->ndim = desc->ndim;
        memcpy(s->shape, desc->shape, sizeof(s->shape));
        s->fail_after = -1;
        s->refs = 1;
        *out = s->h;
        rc = LZYB_OK;
        break;
    }
    pthread_mutex_unlock(&g_mu);
    if (rc != LZYB_OK) {
        munmap(p, desc->nbytes ? desc->nbytes : 1);
        return err(LZYB_E_LOCAL, "too many slots");
    }
    return LZYB_OK;
}

const lzyb_backend_ops_t hostmem_backend_ops = {
    m_slot_create,    m_input_open,     m_slot_commit,         m_slot_read,       m_slot_info,
    m_slot_find,      m_slot_close,     m_transfer_chunk,      m_input_finish,    m_input_write_host,
    m_input_finish_expect, m_input_to_output, m_input_reset, m_slot_retain, m_slot_set_meta,
    m_slot_export, m_peer_import,
};

/* ---- test helpers ---- */
int hostmem_write(lzyb_handle_t h, uint64_t off, const void* src, uint64_t len) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    int rc = -1;
    if (s && off + len <= s->nbytes) {
        memcpy(s->data + off, src, len);
        rc = 0;
    }
    pthread_mutex_unlock(&g_mu);
    return rc;
}
int64_t hostmem_peek(lzyb_handle_t h, uint64_t off, void* dst, uint64_t len) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    int64_t rc = -1;
    if (s && off <= s->nbytes) {
        uint64_t n = s->nbytes - off < len ? s->nbytes - off : len;
        memcpy(dst, s->data + off, n);
        rc = (int64_t)n;
    }
    pthread_mutex_unlock(&g_mu);
    return rc;
}
int hostmem_set_meta(lzyb_handle_t h, uint32_t dtype, uint32_t ndim, const uint64_t* shape) {
    return m_slot_set_meta(h, dtype, ndim, shape);
}
int hostmem_info(lzyb_handle_t h, lzyb_slot_info_t* out) { return m_slot_info(h, out); }
int64_t hostmem_size(lzyb_handle_t h) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    int64_t rc = s ? (int64_t)s->nbytes : -1;
    pthread_mutex_unlock(&g_mu);
    return rc;
}
int hostmem_set_fail_after(lzyb_handle_t h, long long n) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    if (s) s->fail_after = n;
    pthread_mutex_unlock(&g_mu);
    return s ? 0 : -1;
}
int hostmem_set_fail_read(lzyb_handle_t h, int flag) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    if (s) s->fail_read = flag;
    pthread_mutex_unlock(&g_mu);
    return s ? 0 : -1;
}
void hostmem_set_fail_open(int flag) { g_fail_open = flag; }
void hostmem_set_fail_commit(int flag) { g_fail_commit = flag; }
int hostmem_corrupt(lzyb_handle_t h, uint64_t off) {
    pthread_mutex_lock(&g_mu);
    mem_slot_t* s = find(h);
    int rc = -1;
    if (s && off < s->nbytes) {
        s->data[off] ^= 0x5a;
        rc = 0;
    }
    pthread_mutex_unlock(&g_mu);
    return rc;
}
long long hostmem_moving_calls(void) {
    pthread_mutex_lock(&g_mu);
    long long n = g_moving_calls;
    pthread_mutex_unlock(&g_mu);
    return n;
}
int hostmem_live_slots(void) {
    int n = 0;
    pthread_mutex_lock(&g_mu);
    for (int i = 0; i < MAX_SLOTS; ++i) n += g_slots[i].used;
    pthread_mutex_unlock(&g_mu);
    return n;
}
