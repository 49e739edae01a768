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

} // namespace

extern "C" {

char const* agb_op_list() {
    return "comm_alloc,comm_free,comm_ipc_handle,comm_ipc_open,comm_ipc_close,comm_enable_peer,comm_can_access_peer,comm_p2p_copy,comm_barrier_test";
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

} // extern "C"
