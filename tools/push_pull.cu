// push_pull.cu — measurement aid (not part of the product): does a *push* (kernel on the GPU that holds the
// source, stores cross NVLink) beat the shipped *pull* (kernel on the consumer GPU, loads cross NVLink) when
// both directions of a GPU pair are busy?  Uses the product kernels (stream_digest_kernel, tma ring) with
// FIN_NONE so the numbers transfer.  Needs 2 GPUs with peer access.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude -o tools/_build/push_pull tools/push_pull.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <algorithm>

#include "../lzy_b200/csrc/tma_kernels.cuh"

using namespace lzyb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr unsigned long long N = 1ull << 30;
constexpr int REPS = 12;

struct Side {
    int dev;
    uint8_t *a, *b;   // a: source payload, b: sink
    SlotCtl* ctl;
    cudaStream_t st;
    cudaEvent_t e0, e1;
};

enum Engine { LDG2 = 0, LDG_PERSIST = 1, TMA = 2, CE = 3 };

static void launch(Engine eng, const uint8_t* src, uint8_t* dst, SlotCtl* ctl, cudaStream_t st, int sms) {
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = src; p.dst = dst; p.nbytes = N; p.ctl = ctl; p.total_nbytes = N; p.fin = FIN_NONE;
    const unsigned long long tiles = (N >> 4) / (256ull * 8);
    if (eng == LDG2) stream_digest_kernel<16, 8, true><<<(unsigned)(tiles / 2), 256, 0, st>>>(p);
    else if (eng == LDG_PERSIST) stream_digest_kernel<16, 8, true><<<sms * 4, 256, 0, st>>>(p);
    else if (eng == TMA) {
        constexpr size_t smem = 3 * 32768 + 2 * 3 * sizeof(uint64_t);
        tma_stream_digest_kernel<3, 32768, true><<<sms * 2, TMA_THREADS, smem, st>>>(p);
    } else cudaMemcpyAsync(dst, src, N, cudaMemcpyDefault, st);
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("{\"skip\": \"needs 2 GPUs\"}\n"); return 0; }
    Side s[2];
    for (int i = 0; i < 2; ++i) {
        s[i].dev = i;
        CK(cudaSetDevice(i));
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, i, 1 - i));
        if (!can) { printf("{\"skip\": \"no peer access\"}\n"); return 0; }
        CK(cudaDeviceEnablePeerAccess(1 - i, 0));
        CK(cudaMalloc(&s[i].a, N)); CK(cudaMalloc(&s[i].b, N)); CK(cudaMalloc(&s[i].ctl, sizeof(SlotCtl)));
        CK(cudaMemset(s[i].a, 0x5a + i, N)); CK(cudaMemset(s[i].b, 0, N)); CK(cudaMemset(s[i].ctl, 0, sizeof(SlotCtl)));
        CK(cudaStreamCreateWithFlags(&s[i].st, cudaStreamNonBlocking));
        CK(cudaEventCreate(&s[i].e0)); CK(cudaEventCreate(&s[i].e1));
        constexpr size_t smem = 3 * 32768 + 2 * 3 * sizeof(uint64_t);
        CK(cudaFuncSetAttribute(tma_stream_digest_kernel<3, 32768, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

    // mode: pull = the kernel runs on the GPU that owns the sink; push = on the GPU that owns the source.
    // dirs: 1 = only 0 → 1 busy; 2 = 0 → 1 and 1 → 0 at the same time.
    const char* eng_name[] = {"ldg 2 tiles/CTA", "ldg persistent 4/SM", "tma ring 3x32K 2/SM", "copy engine"};
    for (int eng = 0; eng < 4; ++eng)
        for (int push = 0; push < 2; ++push)
            for (int dirs = 1; dirs <= 2; ++dirs) {
                if (eng == CE && push) continue;  // cudaMemcpyAsync has no side
                float ms[2] = {0, 0};
                for (int pass = 0; pass < 2; ++pass) {  // pass 0 = warm-up
                    for (int k = 0; k < dirs; ++k) {
                        // transfer k: source on GPU k, sink on GPU 1-k
                        const int run = push ? k : 1 - k;
                        CK(cudaSetDevice(run));
                        CK(cudaEventRecord(s[run].e0, s[run].st));
                    }
                    for (int r = 0; r < (pass ? REPS : 2); ++r)
                        for (int k = 0; k < dirs; ++k) {
                            const int run = push ? k : 1 - k;
                            CK(cudaSetDevice(run));
                            launch((Engine)eng, s[k].a, s[1 - k].b, s[run].ctl, s[run].st, sms);
                        }
                    for (int k = 0; k < dirs; ++k) {
                        const int run = push ? k : 1 - k;
                        CK(cudaSetDevice(run));
                        CK(cudaEventRecord(s[run].e1, s[run].st));
                    }
                    for (int k = 0; k < dirs; ++k) {
                        const int run = push ? k : 1 - k;
                        CK(cudaSetDevice(run));
                        CK(cudaEventSynchronize(s[run].e1));
                        CK(cudaEventElapsedTime(&ms[k], s[run].e0, s[run].e1));
                    }
                    CK(cudaGetLastError());
                }
                const float worst = dirs == 2 ? std::max(ms[0], ms[1]) : ms[0];
                printf("{\"engine\": \"%s\", \"mode\": \"%s\", \"directions\": %d, \"ms_per_GiB\": %.4f, \"GBps_per_dir\": %.1f}\n",
                       eng_name[eng], push ? "push" : "pull", dirs, worst / REPS, (double)N * REPS / worst / 1e6);
                fflush(stdout);
            }
    // the shipped peer-pull shape: copy engine in sub-chunks on one stream, digest scans of the landed sub-chunks
    // trailing on a second stream of the sink GPU (launch_dma_verify in capi.cu)
    cudaStream_t vs[2], cs[2][4];
    cudaEvent_t evs[2][64];
    for (int i = 0; i < 2; ++i) {
        CK(cudaSetDevice(i));
        CK(cudaStreamCreateWithFlags(&vs[i], cudaStreamNonBlocking));
        for (auto& c : cs[i]) CK(cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking));
        for (auto& slot : evs[i]) CK(cudaEventCreateWithFlags(&slot, cudaEventDisableTiming));
    }
    const unsigned long long subs[] = {32ull << 20, 64ull << 20, 128ull << 20, 256ull << 20, 1024ull << 20};
    for (int ncs = 1; ncs <= 4; ++ncs)
    for (unsigned long long sub : subs)
        for (int dirs = 1; dirs <= 2; ++dirs) {
            if (ncs == 3 || (ncs > 1 && sub > (256ull << 20))) continue;
            float ms[2] = {0, 0};
            for (int pass = 0; pass < 2; ++pass) {
                for (int k = 0; k < dirs; ++k) { CK(cudaSetDevice(1 - k)); CK(cudaEventRecord(s[1 - k].e0, s[1 - k].st)); }
                for (int r = 0; r < (pass ? REPS : 2); ++r)
                    for (int k = 0; k < dirs; ++k) {
                        const int run = 1 - k;  // sink GPU
                        CK(cudaSetDevice(run));
                        int ne = 0, j = 0;
                        if (ncs > 1) {  // the copy streams start after whatever `st` did before
                            cudaEvent_t start = evs[run][ne++ & 63];
                            CK(cudaEventRecord(start, s[run].st));
                            for (int c = 0; c < ncs; ++c) CK(cudaStreamWaitEvent(cs[run][c], start, 0));
                        }
                        for (unsigned long long off = 0; off < N; off += sub, ++j) {
                            const unsigned long long nb = std::min(sub, N - off);
                            cudaStream_t cst = ncs > 1 ? cs[run][j % ncs] : s[run].st;
                            CK(cudaMemcpyAsync(s[run].b + off, s[k].a + off, nb, cudaMemcpyDefault, cst));
                            cudaEvent_t landed = evs[run][ne++ & 63];
                            CK(cudaEventRecord(landed, cst));
                            CK(cudaStreamWaitEvent(vs[run], landed, 0));
                            CopyParams p;
                            memset(&p, 0, sizeof(p));
                            p.src = s[run].b + off; p.dst = nullptr; p.nbytes = nb; p.word_base = off >> 3;
                            p.ctl = s[run].ctl; p.total_nbytes = N; p.fin = FIN_NONE;
                            const unsigned long long tiles = (nb >> 4) / (256ull * 8);
                            stream_digest_kernel<16, 8, false><<<(unsigned)((tiles + 7) / 8), 256, 0, vs[run]>>>(p);
                        }
                        cudaEvent_t verified = evs[run][ne++ & 63];
                        CK(cudaEventRecord(verified, vs[run]));
                        CK(cudaStreamWaitEvent(s[run].st, verified, 0));
                    }
                for (int k = 0; k < dirs; ++k) { CK(cudaSetDevice(1 - k)); CK(cudaEventRecord(s[1 - k].e1, s[1 - k].st)); }
                for (int k = 0; k < dirs; ++k) {
                    CK(cudaSetDevice(1 - k));
                    CK(cudaEventSynchronize(s[1 - k].e1));
                    CK(cudaEventElapsedTime(&ms[k], s[1 - k].e0, s[1 - k].e1));
                }
                CK(cudaGetLastError());
            }
            const float worst = dirs == 2 ? std::max(ms[0], ms[1]) : ms[0];
            printf("{\"engine\": \"copy engine (%d copy streams) + trailing scan, sub-chunk %llu MiB\", \"mode\": \"pull\", \"directions\": %d, "
                   "\"ms_per_GiB\": %.4f, \"GBps_per_dir\": %.1f}\n", ncs, sub >> 20, dirs, worst / REPS, (double)N * REPS / worst / 1e6);
            fflush(stdout);
        }
    // spot check: the sinks hold the peers' bytes
    uint8_t h[2] = {0, 0};
    CK(cudaMemcpy(&h[0], s[1].b + 12345, 1, cudaMemcpyDefault));
    CK(cudaMemcpy(&h[1], s[0].b + (N - 7), 1, cudaMemcpyDefault));
    printf("{\"check\": \"%s\"}\n", (h[0] == 0x5a && h[1] == 0x5b) ? "ok" : "MISMATCH");
    return 0;
}
