// Communication substrate helpers (sm_100a): raw device allocations shareable through CUDA IPC
// (fallback rendezvous when torch's symmetric memory is unavailable), peer-access enabling,
// a P2P read/write bandwidth probe and a cross-GPU flag-barrier self-test.
//
// Replaces the reference's transport layer (gRPC / MPI rendezvous, `tf_patches/`): gradients and
// parameters live in buffers mapped into every peer, kernels dereference peer pointers directly.

#include <agb_device.cuh>

using namespace agb;

namespace {

__global__ void p2p_copy_kernel(float const* __restrict__ src, float* __restrict__ dst, long long n4) {
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (; i < n4; i += stride)
        st_stream_f4(dst + i * 4, ld_stream_f4(src + i * 4));
}

// Every rank: signal all peers (slot `rank` of their pad), wait for all peers, `rounds` times.
__global__ void barrier_test_kernel(unsigned long long const* pads, int R, int rank, uint32_t first_epoch, int rounds) {
    for (int r = 0; r < rounds; ++r) {
        uint32_t epoch = first_epoch + r;
        if (threadIdx.x < R)
            st_release_sys(reinterpret_cast<uint32_t*>(pads[threadIdx.x]) + rank, epoch);
        if (threadIdx.x < R)
            wait_flag_sys(reinterpret_cast<uint32_t*>(pads[rank]) + threadIdx.x, epoch);
        __syncthreads();
    }
}

// ---- collectives over peer-mapped memory ------------------------------------------- //
// The reference ships MPI ring collectives as TF ops (`tf_patches/kernels/ring.h:155-318`, `ring.cu.cc:88-105`: n-1
// send/recv steps + an element-wise accumulate kernel per step). On an NVSwitch box every rank reaches every peer at full
// rate, so both collectives are ONE kernel per rank with no ring: all-reduce = each rank reduces the 1/R slice it owns
// straight from the peers' buffers (in-switch `multimem.ld_reduce` when a multicast mapping exists, P2P loads otherwise) and
// writes the result into every peer (`multimem.st` / P2P stores); all-gather = each rank pulls the peers' blocks.

constexpr int kMaxRanks = 16;

struct CollArgs {
    unsigned long long buf[kMaxRanks];      // the symmetric data buffer as mapped for each rank
    unsigned long long signal[kMaxRanks];   // each rank's signal pad: uint32 [2][R] (entry / exit epochs)
    unsigned long long mc;                  // multicast mapping of the data buffer (0 = none)
    unsigned int* counter;                  // local, zero between launches: blocks that finished their part
    long long offs[kMaxRanks + 1];          // all-gather: block boundaries in 16-byte vectors; all-reduce: offs[1] = total vectors
    unsigned long long out;                 // all-gather: local destination
    int R, rank, mean;
    uint32_t epoch;
};

__device__ __forceinline__ void coll_enter(CollArgs const& a) {
    if (blockIdx.x == 0 && threadIdx.x < a.R)
        st_release_sys(reinterpret_cast<uint32_t*>(a.signal[threadIdx.x]) + a.rank, a.epoch);
    if (threadIdx.x < a.R)
        wait_flag_sys(reinterpret_cast<uint32_t*>(a.signal[a.rank]) + threadIdx.x, a.epoch);
    __syncthreads();
}

// Every block fences its stores; the last block to finish signals the peers and waits for theirs, so that the kernel (hence
// everything after it in the stream) completes only when every peer is done reading from and writing to this rank.
__device__ __forceinline__ void coll_exit(CollArgs const& a) {
    __shared__ int last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last)
        return;
    if (threadIdx.x < a.R)
        st_release_sys(reinterpret_cast<uint32_t*>(a.signal[threadIdx.x]) + a.R + a.rank, a.epoch);
    if (threadIdx.x < a.R)
        wait_flag_sys(reinterpret_cast<uint32_t*>(a.signal[a.rank]) + a.R + threadIdx.x, a.epoch);
    __syncthreads();
    if (threadIdx.x == 0)
        *a.counter = 0;
}

template<typename T> struct Vec16 { static constexpr int kLanes = 16 / sizeof(T); T v[kLanes]; };

template<typename T> __device__ __forceinline__ Vec16<T> ld_vec(unsigned long long base, long long index) {
    Vec16<T> out;
    int4 raw;
    asm volatile("ld.global.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(base + index * 16));
    *reinterpret_cast<int4*>(out.v) = raw;
    return out;
}
template<typename T> __device__ __forceinline__ void st_vec(unsigned long long base, long long index, Vec16<T> const& value) {
    int4 raw = *reinterpret_cast<int4 const*>(value.v);
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1, %2, %3, %4};" :: "l"(base + index * 16), "r"(raw.x), "r"(raw.y), "r"(raw.z), "r"(raw.w) : "memory");
}

template<typename T, bool MC> __global__ void __launch_bounds__(512) allreduce_kernel(CollArgs const a) {
    coll_enter(a);
    long long const total = a.offs[1];
    long long const lo = total * a.rank / a.R, hi = total * (a.rank + 1) / a.R;
    long long const stride = static_cast<long long>(gridDim.x) * blockDim.x;
    if constexpr (MC) {  // fp32 only: the switch adds the R copies and fans the result out; 4 independent round trips per thread
        constexpr int U = 4;
        float const scale = a.mean ? 1.f / static_cast<float>(a.R) : 1.f;
        for (long long v0 = lo + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v0 < hi; v0 += stride * U) {
            float4 sum[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (v0 + u * stride < hi)
                    sum[u] = multimem_ld_reduce_add_f4(reinterpret_cast<float const*>(a.mc + (v0 + u * stride) * 16));
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (v0 + u * stride < hi) {
                    sum[u].x *= scale; sum[u].y *= scale; sum[u].z *= scale; sum[u].w *= scale;
                    multimem_st_f4(reinterpret_cast<float*>(a.mc + (v0 + u * stride) * 16), sum[u]);
                }
        }
    } else {
        for (long long v = lo + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < hi; v += stride) {
            Vec16<T> parts[kMaxRanks > 8 ? 8 : kMaxRanks];
            Vec16<T> sum = ld_vec<T>(a.buf[0], v);  // fixed rank order: every rank would compute the same bits
            for (int r0 = 1; r0 < a.R; r0 += 8) {   // up to 8 peer loads in flight
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (r0 + k < a.R)
                        parts[k] = ld_vec<T>(a.buf[r0 + k], v);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (r0 + k < a.R) {
#pragma unroll
                        for (int i = 0; i < Vec16<T>::kLanes; ++i)
                            sum.v[i] += parts[k].v[i];
                    }
            }
            if (a.mean) {
#pragma unroll
                for (int i = 0; i < Vec16<T>::kLanes; ++i)
                    sum.v[i] = static_cast<T>(sum.v[i] / static_cast<T>(a.R));
            }
            for (int r = 0; r < a.R; ++r)
                st_vec<T>(a.buf[r], v, sum);
        }
    }
    coll_exit(a);
}

__global__ void __launch_bounds__(512) allgather_kernel(CollArgs const a) {
    coll_enter(a);
    long long const total = a.offs[a.R];
    long long const stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
        int r = 0;
        while (v >= a.offs[r + 1])
            ++r;
        st_vec<int>(a.out, v, ld_vec<int>(a.buf[r], v - a.offs[r]));
    }
    coll_exit(a);
}

int fill_args(CollArgs& a, unsigned long long const* ptrs, long long const* offs, int R, int rank, unsigned epoch) {
    if (R < 1 || R > kMaxRanks || rank < 0 || rank >= R)
        return 102;
    for (int r = 0; r < R; ++r) {
        a.buf[r] = ptrs[r];
        a.signal[r] = ptrs[kMaxRanks + r];
    }
    a.mc = ptrs[2 * kMaxRanks];
    a.counter = reinterpret_cast<unsigned int*>(ptrs[2 * kMaxRanks + 1]);
    a.out = ptrs[2 * kMaxRanks + 2];
    for (int r = 0; r <= R; ++r)
        a.offs[r] = offs[r];
    a.R = R;
    a.rank = rank;
    a.epoch = epoch;
    return 0;
}

int grid_for(long long vectors, int max_blocks) {
    long long want = (vectors + 511) / 512;
    int cap = max_blocks > 0 ? max_blocks : 148 * 2;
    return static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
}

} // namespace

extern "C" {

char const* agb_op_list() {
    return "comm_alloc,comm_free,comm_ipc_handle,comm_ipc_open,comm_ipc_close,comm_enable_peer,comm_can_access_peer,comm_p2p_copy,comm_barrier_test,comm_allreduce,comm_allgather";
}

// Wall-clock bound (seconds, 0 = none) of the cross-GPU flag waits of this library's kernels.
int agb_comm_set_flag_timeout(double seconds) {
    return set_flag_timeout(seconds);
}

int agb_comm_alloc(unsigned long long size, unsigned long long* out) {
    void* ptr = nullptr;
    AGB_CUDA_OK(cudaMalloc(&ptr, size));
    AGB_CUDA_OK(cudaMemset(ptr, 0, size));
    *out = reinterpret_cast<unsigned long long>(ptr);
    return 0;
}

int agb_comm_free(unsigned long long ptr) {
    AGB_CUDA_OK(cudaFree(reinterpret_cast<void*>(ptr)));
    return 0;
}

int agb_comm_ipc_handle(unsigned long long ptr, unsigned char* out64) {
    cudaIpcMemHandle_t handle;
    AGB_CUDA_OK(cudaIpcGetMemHandle(&handle, reinterpret_cast<void*>(ptr)));
    static_assert(sizeof(handle) == 64, "unexpected IPC handle size");
    for (int i = 0; i < 64; ++i)
        out64[i] = reinterpret_cast<unsigned char const*>(&handle)[i];
    return 0;
}

int agb_comm_ipc_open(unsigned char const* in64, unsigned long long* out) {
    cudaIpcMemHandle_t handle;
    for (int i = 0; i < 64; ++i)
        reinterpret_cast<unsigned char*>(&handle)[i] = in64[i];
    void* ptr = nullptr;
    AGB_CUDA_OK(cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess));
    *out = reinterpret_cast<unsigned long long>(ptr);
    return 0;
}

int agb_comm_ipc_close(unsigned long long ptr) {
    AGB_CUDA_OK(cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr)));
    return 0;
}

int agb_comm_can_access_peer(int device, int peer) {
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, device, peer) != cudaSuccess)
        return 0;
    return can;
}

int agb_comm_enable_peer(int peer) {
    cudaError_t err = cudaDeviceEnablePeerAccess(peer, 0);
    if (err == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return 0;
    }
    AGB_CUDA_OK(err);
    return 0;
}

int agb_comm_p2p_copy(unsigned long long src, unsigned long long dst, long long nfloats, int blocks, void* stream) {
    if (nfloats & 3)
        return 101;
    p2p_copy_kernel<<<blocks > 0 ? blocks : 148 * 4, 512, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float const*>(src), reinterpret_cast<float*>(dst), nfloats / 4);
    AGB_CUDA_OK(cudaGetLastError());
    return 0;
}

int agb_comm_barrier_test(unsigned long long const* pads_dev, int R, int rank, unsigned first_epoch, int rounds, void* stream) {
    barrier_test_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pads_dev, R, rank, first_epoch, rounds);
    AGB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ptrs: [0,16) data buffer per rank, [16,32) signal pad per rank, [32] multicast data pointer or 0, [33] local counter, [34] all-gather output.
// dtype: 0 = fp32, 1 = int32, 2 = int64. `vectors` = 16-byte vectors in the buffer (the same on every rank).
int agb_comm_allreduce(unsigned long long const* ptrs, long long vectors, int dtype, int mean, int R, int rank, unsigned epoch, int max_blocks, void* stream) {
    CollArgs a{};
    long long offs[kMaxRanks + 1] = {0, vectors};  // only offs[1] is read by the all-reduce
    int status = fill_args(a, ptrs, offs, R, rank, epoch);
    if (status)
        return status;
    a.offs[1] = vectors;
    a.mean = mean;
    int const blocks = grid_for((vectors + R - 1) / R, max_blocks);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == 0 && a.mc != 0 && R > 1)
        allreduce_kernel<float, true><<<blocks, 512, 0, s>>>(a);
    else if (dtype == 0)
        allreduce_kernel<float, false><<<blocks, 512, 0, s>>>(a);
    else if (dtype == 1)
        allreduce_kernel<int, false><<<blocks, 512, 0, s>>>(a);
    else if (dtype == 2)
        allreduce_kernel<long long, false><<<blocks, 512, 0, s>>>(a);
    else
        return 103;
    AGB_CUDA_OK(cudaGetLastError());
    return 0;
}

// offs[r] .. offs[r + 1]: where rank r's block lands in the output, in 16-byte vectors (variable block sizes allowed).
int agb_comm_allgather(unsigned long long const* ptrs, long long const* offs, int R, int rank, unsigned epoch, int max_blocks, void* stream) {
    CollArgs a{};
    int status = fill_args(a, ptrs, offs, R, rank, epoch);
    if (status)
        return status;
    allgather_kernel<<<grid_for(offs[R], max_blocks), 512, 0, static_cast<cudaStream_t>(stream)>>>(a);
    AGB_CUDA_OK(cudaGetLastError());
    return 0;
}

} // extern "C"
