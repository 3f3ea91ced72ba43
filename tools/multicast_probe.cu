// multicast_probe.cu — next-round measurement aid (not part of the product): can one stream of NVSwitch
// multicast stores (`multimem.st` into a cuMulticast object bound on every consumer GPU) deliver a 1 → R fan-out
// faster than NCCL's broadcast (6.6 ms for 4 GiB, 1 → 7) and the device-chained pull kernel (8.1 ms)?  The bound
// is the root's NVLink egress: 4 GiB / 0.77 TB/s ≈ 5.6 ms regardless of R.
//
// Single process, all GPUs of the box.  For every team size R+1 in {2, 4, 8} (as available):
//   * physical memory on each GPU (cuMemCreate), mapped locally (unicast VA) and bound into one multicast object;
//   * the root runs a copy kernel that reads its source with LDG.128 and writes the *multicast* VA with
//     multimem.st.v4 — the switch replicates every store to all bound GPUs (the root's own copy included);
//   * each receiver then checks a sample of its unicast mapping against the source pattern.
// Reported: ms per 1 GiB payload and delivered GB/s per receiver, next to R unicast cudaMemcpyPeerAsync copies.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/multicast_probe tools/multicast_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#define DRV(x)                                                                         \
    do {                                                                               \
        CUresult r_ = (x);                                                             \
        if (r_ != CUDA_SUCCESS) {                                                      \
            const char* s_ = nullptr;                                                  \
            cuGetErrorString(r_, &s_);                                                 \
            printf("{\"error\": \"%s at line %d: %s\"}\n", #x, __LINE__, s_ ? s_ : "?"); \
            return 1;                                                                  \
        }                                                                              \
    } while (0)
#define RT(x)                                                                                        \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) {                                                                     \
            printf("{\"error\": \"%s at line %d: %s\"}\n", #x, __LINE__, cudaGetErrorString(e_));    \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

__device__ __forceinline__ void multimem_st_v4(void* mc_addr, const uint4& v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}

// one 32 KiB tile per 256-thread CTA, 8 vectors in flight per thread (the product's LDG engine shape)
__global__ void __launch_bounds__(256) multicast_copy(const uint4* __restrict__ src, uint4* mc_dst, unsigned long long nvec) {
    const unsigned long long base = (unsigned long long)blockIdx.x * 2048ull + threadIdx.x;
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (base + j * 256ull < nvec) v[j] = src[base + j * 256ull];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (base + j * 256ull < nvec) multimem_st_v4(mc_dst + base + j * 256ull, v[j]);
}

// persistent variant: 4 CTAs per SM walk the tiles grid-stride (fewer, longer-lived store streams)
__global__ void __launch_bounds__(256) multicast_copy_persistent(const uint4* __restrict__ src, uint4* mc_dst, unsigned long long nvec) {
    const unsigned long long ntiles = (nvec + 2047) / 2048;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned long long base = t * 2048ull + threadIdx.x;
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (base + j * 256ull < nvec) v[j] = src[base + j * 256ull];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (base + j * 256ull < nvec) multimem_st_v4(mc_dst + base + j * 256ull, v[j]);
    }
}

__global__ void fill(uint4* p, unsigned long long nvec) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nvec; i += (unsigned long long)gridDim.x * blockDim.x)
        p[i] = make_uint4((unsigned)i, (unsigned)(i >> 32) ^ 0x9e3779b9u, (unsigned)(i * 2654435761ull), 0x7fc00001u /* a NaN pattern: bits must survive */);
}

__global__ void check(const uint4* p, unsigned long long nvec, unsigned long long stride, unsigned long long* bad) {
    for (unsigned long long i = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * stride; i < nvec;
         i += (unsigned long long)gridDim.x * blockDim.x * stride) {
        const uint4 v = p[i];
        if (v.x != (unsigned)i || v.y != ((unsigned)(i >> 32) ^ 0x9e3779b9u) || v.z != (unsigned)(i * 2654435761ull) || v.w != 0x7fc00001u)
            atomicAdd(bad, 1ull);
    }
}

int main() {
    DRV(cuInit(0));
    int ndev = 0;
    RT(cudaGetDeviceCount(&ndev));
    if (ndev < 2) {
        printf("{\"skip\": \"needs at least 2 GPUs\"}\n");
        return 0;
    }
    for (int d = 0; d < ndev; ++d) {
        int ok = 0;
        CUdevice dev;
        DRV(cuDeviceGet(&dev, d));
        DRV(cuDeviceGetAttribute(&ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
        if (!ok) {
            printf("{\"skip\": \"device %d does not support multicast\"}\n", d);
            return 0;
        }
        RT(cudaSetDevice(d));
        RT(cudaFree(nullptr));  // primary context
        for (int p = 0; p < ndev; ++p)
            if (p != d) cudaDeviceEnablePeerAccess(p, 0);
    }
    const size_t payload = 1ull << 30;
    const unsigned long long nvec = payload / 16;

    for (int team : {2, 4, 8}) {
        if (team > ndev) break;
        CUmulticastObjectProp mprop = {};
        mprop.numDevices = (unsigned)team;
        mprop.size = payload;
        mprop.handleTypes = CU_MEM_HANDLE_TYPE_NONE;
        size_t gran = 0;
        DRV(cuMulticastGetGranularity(&gran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
        const size_t size = (payload + gran - 1) / gran * gran;
        mprop.size = size;
        CUmemGenericAllocationHandle mc;
        DRV(cuMulticastCreate(&mc, &mprop));
        for (int d = 0; d < team; ++d) {
            CUdevice dev;
            DRV(cuDeviceGet(&dev, d));
            DRV(cuMulticastAddDevice(mc, dev));
        }
        std::vector<CUmemGenericAllocationHandle> mem(team);
        std::vector<CUdeviceptr> uc(team);
        for (int d = 0; d < team; ++d) {
            RT(cudaSetDevice(d));
            CUmemAllocationProp ap = {};
            ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
            ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            ap.location.id = d;
            DRV(cuMemCreate(&mem[d], size, &ap, 0));
            DRV(cuMulticastBindMem(mc, 0, mem[d], 0, size, 0));
            DRV(cuMemAddressReserve(&uc[d], size, gran, 0, 0));
            DRV(cuMemMap(uc[d], size, 0, mem[d], 0));
            CUmemAccessDesc ad = {};
            ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            ad.location.id = d;
            ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            DRV(cuMemSetAccess(uc[d], size, &ad, 1));
            RT(cudaMemset((void*)uc[d], 0, payload));
        }
        // the root writes through the multicast mapping
        RT(cudaSetDevice(0));
        CUdeviceptr mcva;
        DRV(cuMemAddressReserve(&mcva, size, gran, 0, 0));
        DRV(cuMemMap(mcva, size, 0, mc, 0));
        CUmemAccessDesc rootad = {};
        rootad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        rootad.location.id = 0;
        rootad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        DRV(cuMemSetAccess(mcva, size, &rootad, 1));
        uint4* src = nullptr;
        RT(cudaMalloc(&src, payload));
        fill<<<148 * 8, 256>>>(src, nvec);
        RT(cudaDeviceSynchronize());

        cudaEvent_t e0, e1;
        RT(cudaEventCreate(&e0));
        RT(cudaEventCreate(&e1));
        const unsigned grid = (unsigned)((nvec + 2047) / 2048);
        const int reps = 10;
        for (int pass = 0; pass < 2; ++pass) {
            RT(cudaEventRecord(e0));
            for (int r = 0; r < (pass ? reps : 2); ++r) multicast_copy<<<grid, 256>>>(src, (uint4*)mcva, nvec);
            RT(cudaEventRecord(e1));
            RT(cudaEventSynchronize(e1));
        }
        RT(cudaGetLastError());
        float ms = 0;
        RT(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;

        // other carriers of the same multicast stream: a persistent SM kernel, and the copy engine writing the
        // multicast mapping directly
        float ms_persist = 0, ms_ce = 0;
        for (int pass = 0; pass < 2; ++pass) {
            RT(cudaEventRecord(e0));
            for (int r = 0; r < (pass ? reps : 2); ++r) multicast_copy_persistent<<<148 * 4, 256>>>(src, (uint4*)mcva, nvec);
            RT(cudaEventRecord(e1));
            RT(cudaEventSynchronize(e1));
        }
        RT(cudaGetLastError());
        RT(cudaEventElapsedTime(&ms_persist, e0, e1));
        ms_persist /= reps;
        bool ce_ok = true;
        for (int pass = 0; pass < 2 && ce_ok; ++pass) {
            RT(cudaEventRecord(e0));
            for (int r = 0; r < (pass ? reps : 2); ++r)
                if (cudaMemcpyAsync((void*)mcva, src, payload, cudaMemcpyDeviceToDevice) != cudaSuccess) ce_ok = false;
            RT(cudaEventRecord(e1));
            if (cudaEventSynchronize(e1) != cudaSuccess) ce_ok = false;
        }
        if (ce_ok) {
            RT(cudaEventElapsedTime(&ms_ce, e0, e1));
            ms_ce /= reps;
        } else {
            cudaGetLastError();
        }

        unsigned long long bad_total = 0;
        for (int d = 0; d < team; ++d) {
            RT(cudaSetDevice(d));
            unsigned long long* bad = nullptr;
            RT(cudaMalloc(&bad, 8));
            RT(cudaMemset(bad, 0, 8));
            check<<<148, 256>>>((const uint4*)uc[d], nvec, 97, bad);
            unsigned long long h = 0;
            RT(cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost));
            bad_total += h;
            RT(cudaFree(bad));
        }

        // the same delivery as R unicast peer copies from the root, issued back to back on R streams
        RT(cudaSetDevice(0));
        std::vector<cudaStream_t> st(team);
        for (int d = 1; d < team; ++d) RT(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
        float ums = 0;
        for (int pass = 0; pass < 2; ++pass) {
            RT(cudaDeviceSynchronize());
            RT(cudaEventRecord(e0));
            for (int r = 0; r < (pass ? reps : 1); ++r)
                for (int d = 1; d < team; ++d) RT(cudaMemcpyPeerAsync((void*)uc[d], d, src, 0, payload, st[d]));
            for (int d = 1; d < team; ++d) RT(cudaStreamSynchronize(st[d]));
            RT(cudaEventRecord(e1));
            RT(cudaEventSynchronize(e1));
        }
        RT(cudaEventElapsedTime(&ums, e0, e1));
        ums /= reps;

        printf("{\"team\": %d, \"receivers\": %d, \"payload_GiB\": 1, \"multicast_ms\": %.4f, \"multicast_GBps_per_receiver\": %.1f, "
               "\"multicast_persistent_ms\": %.4f, \"multicast_copy_engine_ms\": %.4f, \"unicasts_ms\": %.4f, "
               "\"mismatching_samples\": %llu, \"granularity\": %zu}\n",
               team, team - 1, ms, payload / ms / 1e6, ms_persist, ms_ce, ums, bad_total, gran);
        fflush(stdout);

        for (int d = 1; d < team; ++d) cudaStreamDestroy(st[d]);
        cudaFree(src);
        cuMemUnmap(mcva, size);
        cuMemAddressFree(mcva, size);
        for (int d = 0; d < team; ++d) {
            CUdevice dev;
            cuDeviceGet(&dev, d);
            cuMulticastUnbind(mc, dev, 0, size);
            cuMemUnmap(uc[d], size);
            cuMemAddressFree(uc[d], size);
            cuMemRelease(mem[d]);
        }
        cuMemRelease(mc);
    }
    return 0;
}
