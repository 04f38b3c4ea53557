// oc_exec.h — kernel launch abstraction.
//
// Every kernel of this library is a POD functor with
//     static constexpr int PHASES;                       // number of __syncthreads()-separated phases
//     template <int PHASE> OC_HD void run(const oc::Block& b, int tid, int nthreads, char* smem) const;
// Phases communicate through shared memory only.  On the GPU (the product build, nvcc, sm_100a) the
// phases run inside one __global__ kernel separated by __syncthreads().  Under OC_HOSTSIM (a
// TEST-ONLY build with g++, see tests/hostsim/README.md) the same functor is executed block by block,
// phase by phase, thread by thread on the host so that index arithmetic and physics can be
// debugged in a container without a GPU.  The host simulation is never built into, nor loaded by,
// the product library: liboceananigans_b200.so has no CPU path.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#ifdef OC_HOSTSIM
#define OC_HD inline
#define OC_DEV inline
#define OC_RESTRICT
#include <cmath>
#else
#include <cuda_runtime.h>
#define OC_HD __host__ __device__ __forceinline__
#define OC_DEV __device__ __forceinline__
#define OC_RESTRICT __restrict__
#endif

namespace oc {

struct Block {
    int x, y, z;     // block index
};

struct Dim3 {
    int x = 1, y = 1, z = 1;
};

#ifndef OC_HOSTSIM
typedef cudaStream_t Stream;
enum { OC_MAX_DEVICES = 64 };

template <class K>
__global__ void __launch_bounds__(K::THREADS, K::MIN_BLOCKS) kernel_entry(const __grid_constant__ K k) {
    extern __shared__ __align__(128) char smem[];
    Block b{(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z};
    k.template run<0>(b, (int)threadIdx.x, (int)blockDim.x, smem);
    if constexpr (K::PHASES > 1) {
        __syncthreads();
        k.template run<1>(b, (int)threadIdx.x, (int)blockDim.x, smem);
    }
    if constexpr (K::PHASES > 2) {
        __syncthreads();
        k.template run<2>(b, (int)threadIdx.x, (int)blockDim.x, smem);
    }
    if constexpr (K::PHASES > 3) {
        __syncthreads();
        k.template run<3>(b, (int)threadIdx.x, (int)blockDim.x, smem);
    }
    static_assert(K::PHASES <= 4, "add more phases to kernel_entry");
}

template <class K>
inline cudaError_t launch(const K& k, Dim3 grid, size_t smem_bytes, Stream stream) {
    if (grid.x <= 0 || grid.y <= 0 || grid.z <= 0) return cudaSuccess;
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: cache what was configured per device
    static size_t configured[OC_MAX_DEVICES] = {};
    if (smem_bytes > 48 * 1024) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= OC_MAX_DEVICES || configured[dev] < smem_bytes) {
            e = cudaFuncSetAttribute(kernel_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < OC_MAX_DEVICES) configured[dev] = smem_bytes;
        }
    }
    kernel_entry<K><<<dim3(grid.x, grid.y, grid.z), K::THREADS, smem_bytes, stream>>>(k);
    return cudaGetLastError();
}

// ---- iterated kernels (z-marching) -------------------------------------------------------------------------
// A marching kernel is a POD functor with
//     OC_DEV void begin0(b, tid, smem) const;            // mbarrier init
//     OC_DEV void begin1(b, tid, smem, State&) const;    // prologue loads, per-thread marching state
//     OC_HD  int  iterations(b) const;                   // number of plane iterations of this block
//     template <int PHASE> OC_DEV void step(b, tid, smem, it, State&) const;   // PHASE 0, 1, 2 for it = 0 … iterations()
//     OC_DEV void sync_wait(smem, it) / sync_arrive(smem, it) const;            // split (arrive / wait) block barrier
template <class K>
__global__ void __launch_bounds__(K::THREADS, K::MIN_BLOCKS) march_entry(const __grid_constant__ K k) {
    extern __shared__ __align__(128) char smem[];
    Block b{(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z};
    const int tid = (int)threadIdx.x;
    typename K::State st;
    k.begin0(b, tid, smem);
    __syncthreads();
    k.begin1(b, tid, smem, st);
    const int n = k.iterations(b);
#pragma unroll 1
    for (int it = 0; it <= n; ++it) {
        k.template step<0>(b, tid, smem, it, st);
        if (it >= 1) k.sync_wait(smem, it - 1);          // split barrier: every thread has arrived for iteration it-1
        k.template step<1>(b, tid, smem, it, st);
        k.template step<2>(b, tid, smem, it, st);
        if (it < n) k.sync_arrive(smem, it);
    }
}

template <class K>
inline cudaError_t launch_march(const K& k, Dim3 grid, size_t smem_bytes, Stream stream) {
    if (grid.x <= 0 || grid.y <= 0 || grid.z <= 0) return cudaSuccess;
    static size_t configured[OC_MAX_DEVICES] = {};
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= OC_MAX_DEVICES || configured[dev] < smem_bytes) {
            e = cudaFuncSetAttribute(march_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < OC_MAX_DEVICES) configured[dev] = smem_bytes;
        }
    }
    march_entry<K><<<dim3(grid.x, grid.y, grid.z), K::THREADS, smem_bytes, stream>>>(k);
    return cudaGetLastError();
}
#else
typedef void* Stream;
typedef int cudaError_t;
static const int cudaSuccess = 0;

template <class K, int PHASE>
inline void hostsim_phase(const K& k, const Block& b, char* smem) {
    for (int tid = 0; tid < K::THREADS; ++tid) k.template run<PHASE>(b, tid, K::THREADS, smem);
}

template <class K>
inline cudaError_t launch(const K& k, Dim3 grid, size_t smem_bytes, Stream) {
    std::vector<char> smem(smem_bytes + 128);
    char* sm = smem.data();
    sm += (128 - ((uintptr_t)sm & 127)) & 127;
    for (int bz = 0; bz < grid.z; ++bz)
        for (int by = 0; by < grid.y; ++by)
            for (int bx = 0; bx < grid.x; ++bx) {
                Block b{bx, by, bz};
                hostsim_phase<K, 0>(k, b, sm);
                if constexpr (K::PHASES > 1) hostsim_phase<K, 1>(k, b, sm);
                if constexpr (K::PHASES > 2) hostsim_phase<K, 2>(k, b, sm);
                if constexpr (K::PHASES > 3) hostsim_phase<K, 3>(k, b, sm);
            }
    return cudaSuccess;
}

template <class K>
inline cudaError_t launch_march(const K& k, Dim3 grid, size_t smem_bytes, Stream) {
    std::vector<char> smem(smem_bytes + 128);
    char* sm = smem.data();
    sm += (128 - ((uintptr_t)sm & 127)) & 127;
    for (int bz = 0; bz < grid.z; ++bz)
        for (int by = 0; by < grid.y; ++by)
            for (int bx = 0; bx < grid.x; ++bx) {
                Block b{bx, by, bz};
                std::vector<typename K::State> st(K::THREADS);
                for (int tid = 0; tid < K::THREADS; ++tid) k.begin0(b, tid, sm);
                for (int tid = 0; tid < K::THREADS; ++tid) k.begin1(b, tid, sm, st[tid]);
                const int n = k.iterations(b);
                for (int it = 0; it <= n; ++it) {
                    for (int tid = 0; tid < K::THREADS; ++tid) k.template step<0>(b, tid, sm, it, st[tid]);
                    for (int tid = 0; tid < K::THREADS; ++tid) k.template step<1>(b, tid, sm, it, st[tid]);
                    for (int tid = 0; tid < K::THREADS; ++tid) k.template step<2>(b, tid, sm, it, st[tid]);
                }
            }
    return cudaSuccess;
}
#endif

}  // namespace oc
