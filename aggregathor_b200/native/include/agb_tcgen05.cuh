// Blackwell (sm_100a) primitives used by the GEMM / implicit-GEMM kernels: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences) and the UMMA shared-memory
// and instruction descriptors. Inline PTX only; bit layouts follow the PTX ISA (cross-checked against
// cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace agb {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(void const* ptr) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------- //
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// Bounded wait (2 s of wall clock): a pipeline bug becomes a trap with a message instead of a hung GPU.
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
    if (mbar_try_wait(bar, parity))
        return;
    uint64_t const start = global_timer_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (global_timer_ns() - start > 2000000000ull) {
            printf("[agb] mbarrier timeout tag %d block (%d,%d,%d) thread %d parity %u\n", tag, (int) blockIdx.x, (int) blockIdx.y, (int) blockIdx.z, (int) threadIdx.x, parity);
            __trap();
        }
    }
}

// ---- TMA -------------------------------------------------------------------- //
__device__ __forceinline__ void tma_prefetch_desc(CUtensorMap const* map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, CUtensorMap const* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, CUtensorMap const* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, CUtensorMap const* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------ //
template<uint32_t kCols> __device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template<uint32_t kCols> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16/fp16 inputs, fp32 accumulation.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(static_cast<uint32_t>(accumulate)) : "memory");
}
// Same with fp32 operands read as TF32 (10-bit mantissa), fp32 accumulation.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(static_cast<uint32_t>(accumulate)) : "memory");
}
// Arrive on `bar` once every previously issued tcgen05.mma of this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t of the warp = lane base + t).
// The asynchronous form lets several loads be in flight before one tcgen05.wait::ld.
__device__ __forceinline__ void tmem_ld_32x32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    tmem_ld_32x32_async(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
        v[i] = __uint_as_float(r[i]);
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of a TPC issue ONE MMA over a 256-row tile ------------- //
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    return rank;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_addr), "r"(rank));
    return out;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" :: "r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to a barrier of the pair's leader CTA (`bar_cluster_addr` from mapa_shared(.., 0))
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, CUtensorMap const* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
template<uint32_t kCols> __device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {   // one warp of EACH CTA of the pair, same dst offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template<uint32_t kCols> __device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA's smem] * B[N columns: N/2 from each CTA's smem]; issued by the leader CTA only.
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(static_cast<uint32_t>(accumulate)) : "memory");
}
// Arrive on the barrier at this shared-memory offset in every CTA of `cta_mask` once the MMAs issued so far have completed.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" :: "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// ---- descriptors ---------------------------------------------------------------- //
// Shared-memory matrix descriptor, 128-byte swizzle, bf16/fp16 operands.
//   K-major  (rows = M/N index, 64 K-elements = 128 B per row): SBO = 1024 B (8 rows), LBO unused (1).
//   MN-major (rows = K index, 64 MN-elements = 128 B per row): SBO = 1024 B (8 k-rows), LBO = byte distance
//            between consecutive 64-element MN chunks.
// `layout_type`: 2 = SWIZZLE_128B (16-byte chunks; every bf16 operand, K-major tf32 operands), 1 = SWIZZLE_128B_BASE32B (32-byte chunks
// over 4-row atoms: the only layout of MN-major 32-bit operands; its k-groups are 4 rows = 512 B apart when rows are packed).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);            // start address, bits [0,14)
    desc |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;       // leading byte offset, bits [16,30)
    desc |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;       // stride byte offset, bits [32,46)
    desc |= static_cast<uint64_t>(1) << 46;                               // descriptor version (Blackwell)
    desc |= static_cast<uint64_t>(layout_type) << 61;                     // layout type
    return desc;
}
// Instruction descriptor for kind::f16 / kind::tf32 with fp32 accumulators; `format`: 0 = F16, 1 = BF16, 2 = TF32 (both operands).
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n, bool a_mn_major, bool b_mn_major, uint32_t format) {
    return (1u << 4)                                   // D format: F32
         | (format << 7)                               // A format
         | (format << 10)                              // B format
         | (static_cast<uint32_t>(a_mn_major) << 15)
         | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(n >> 3) << 17)
         | (static_cast<uint32_t>(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                                   // D format: F32
         | (1u << 7)                                   // A format: BF16
         | (1u << 10)                                  // B format: BF16
         | (static_cast<uint32_t>(a_mn_major) << 15)
         | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(n >> 3) << 17)
         | (static_cast<uint32_t>(m >> 4) << 24);
}

} // namespace sm100

// ---- host: tensor-map creation through the driver entry point (no link-time libcuda dependency) ---- //
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, cuuint64_t const*, cuuint64_t const*, cuuint32_t const*,
                                  cuuint32_t const*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) != cudaSuccess || status != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// Tensor maps are pure functions of (base, dims, strides, box, element strides): the same layer is launched with the same arguments every
// step, so encoded descriptors are memoised (cuTensorMapEncodeTiled is a driver call of a few microseconds; an eager step — evaluation,
// tracing, experiments that opt out of CUDA graphs — made ~300 of them).
struct TmapKey {
    void const* base;
    uint64_t dims[4], strides[3];
    uint32_t box[4], elem[4], rank, dtype, swizzle, pad;
    bool operator==(TmapKey const& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(TmapKey const& k) const {
        uint64_t h = 1469598103934665603ull;
        unsigned char const* p = reinterpret_cast<unsigned char const*>(&k);
        for (size_t i = 0; i < sizeof(TmapKey); ++i)
            h = (h ^ p[i]) * 1099511628211ull;
        return static_cast<size_t>(h);
    }
};
inline int encode_cached(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, void const* base, cuuint64_t const* dims, cuuint64_t const* strides, cuuint32_t const* box,
                         cuuint32_t const* elem, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    static std::mutex mutex;
    TmapKey key;
    std::memset(&key, 0, sizeof(key));
    key.base = base; key.rank = rank; key.dtype = static_cast<uint32_t>(dtype); key.swizzle = static_cast<uint32_t>(swizzle);
    for (uint32_t i = 0; i < rank; ++i) {
        key.dims[i] = dims[i]; key.box[i] = box[i]; key.elem[i] = elem[i];
        if (i + 1 < rank)
            key.strides[i] = strides[i];
    }
    {
        std::lock_guard<std::mutex> guard(mutex);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *map = it->second;
            return 0;
        }
    }
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn)
        return 201;
    CUresult res = fn(map, dtype, rank, const_cast<void*>(base), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS) {
        std::fprintf(stderr, "[agb] cuTensorMapEncodeTiled(rank %u) failed (%d): base %p dims %llu %llu %llu %llu box %u %u %u %u\n", rank, (int) res, base, (unsigned long long) dims[0],
                     (unsigned long long) (rank > 1 ? dims[1] : 0), (unsigned long long) (rank > 2 ? dims[2] : 0), (unsigned long long) (rank > 3 ? dims[3] : 0), box[0],
                     rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return 202;
    }
    std::lock_guard<std::mutex> guard(mutex);
    if (cache.size() > 16384)   // addresses change when buffers are re-allocated: keep the table bounded
        cache.clear();
    cache.emplace(key, *map);
    return 0;
}

// 2-D bf16 tensor map: `inner` contiguous elements per row, `rows` rows of `row_stride` elements, box = box_inner x box_rows, 128B swizzle.
// MN-major 32-bit (tf32) operands must be laid out with 32-byte swizzle chunks (UMMA SWIZZLE_128B_BASE32B); everything else with 16-byte chunks.
inline CUtensorMapSwizzle tmap_swizzle(int elem_bytes, bool mn_major) {
    return (elem_bytes == 4 && mn_major) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
}
inline int make_tmap_2d(CUtensorMap* map, void const* base, uint64_t inner, uint64_t rows, uint64_t row_stride, uint32_t box_inner, uint32_t box_rows, int elem_bytes, bool mn_major = false) {
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_stride * elem_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t elem[2] = {1, 1};
    return encode_cached(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, elem, tmap_swizzle(elem_bytes, mn_major));
}
inline int make_tmap_2d_bf16(CUtensorMap* map, void const* base, uint64_t inner, uint64_t rows, uint64_t row_stride, uint32_t box_inner, uint32_t box_rows) {
    return make_tmap_2d(map, base, inner, rows, row_stride, box_inner, box_rows, 2);
}

// 3-D bf16 tensor map over a row-major [groups][rows][inner] view (row stride `row_stride`, group stride `rows * row_stride`):
// boxes never straddle a group, rows past the end of a group are zero-filled (grouped weight-gradient GEMMs).
inline int make_tmap_3d(CUtensorMap* map, void const* base, uint64_t inner, uint64_t rows, uint64_t groups, uint64_t row_stride, uint32_t box_inner, uint32_t box_rows, int elem_bytes, bool mn_major = false) {
    cuuint64_t dims[3] = {inner, rows, groups};
    cuuint64_t strides[2] = {row_stride * elem_bytes, rows * row_stride * elem_bytes};
    cuuint32_t box[3] = {box_inner, box_rows, 1};
    cuuint32_t elem[3] = {1, 1, 1};
    return encode_cached(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, elem, tmap_swizzle(elem_bytes, mn_major));
}
inline int make_tmap_3d_bf16(CUtensorMap* map, void const* base, uint64_t inner, uint64_t rows, uint64_t groups, uint64_t row_stride, uint32_t box_inner, uint32_t box_rows) {
    return make_tmap_3d(map, base, inner, rows, groups, row_stride, box_inner, box_rows, 2);
}

} // namespace agb
