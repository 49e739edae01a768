#pragma once
// bf16 GEMM on CTA pairs: C[M,N] = A[M,K] * B[N,K]^T (both operands K-major: forward of dense layers / 1x1 convolutions), one
// 256 x 256 output tile per pair of CTAs sitting on the two SMs of a TPC (`tcgen05.mma.cta_group::2`).
//
// Why pairs: a single-CTA 128 x 128 x 16 MMA reads 8 KB of shared memory per 64 tensor-pipe cycles — exactly the 128 B/cycle an SM's
// shared memory delivers, so the tensor pipe starves (47 % active in `profiles/ncu_gemm.txt`). In a pair each SM stages only ITS
// half of both operands (128 rows of A, 128 of the 256 rows of B): the same 8 KB now feed a 256 x 256 x 16 MMA (128 cycles), half
// the shared-memory bandwidth and half the L2 traffic per flop.
//
// Per CTA (both CTAs run the same program, `rank` = %cluster_ctarank):
//   warp 0     TMA producer of its halves; completion bytes are credited to the LEADER's `full` barrier (cta_group::2 loads), which
//              the leader alone arms with the byte count of both CTAs
//   warp 1     TMEM allocation (cta_group::2, both CTAs); in the leader only: one lane issues the MMAs, `tcgen05.commit` multicast
//              releases the ring slot in both CTAs and finally arms both CTAs' `tmem_full`
//   warps 2-9  epilogue of this CTA's 128 rows (TMEM lanes are per CTA), then a remote arrive on the leader's `tmem_empty`
// Persistent: pairs loop over tiles; the accumulator is double-buffered (2 x 256 TMEM columns).

#include "gemm_kernels.cuh"

namespace {

template<int BN> struct PairConfig {
    static constexpr int kStages = 5;
    static constexpr uint32_t kABytes = kBM * 128;            // this CTA's 128 rows of A
    static constexpr uint32_t kBBytes = (BN / 2) * 128;       // this CTA's half of the B rows
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 + 256 + kEpilogueWarps * kStageBytesPerWarp;   // ring + barriers + epilogue staging
    static constexpr uint32_t kTmemCols = 2 * BN;             // double-buffered accumulator: 512 columns for BN = 256
};

template<int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPersistentThreads, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, GemmParams const p, int m_tiles, int n_tiles) {
    using Cfg = PairConfig<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty = full + Cfg::kStages;
    uint64_t* tmem_full = empty + Cfg::kStages;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2] (the leader's are used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint8_t* epi_stage = smem + Cfg::kStages * Cfg::kStageBytes + 256;

    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t const rank = cluster_ctarank();
    bool const leader = rank == 0;
    int const total_kblocks = (p.K + kBK - 1) / kBK;
    int const total_items = m_tiles * n_tiles;
    int const pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(full + s, 1);      // the leader's arrive.expect_tx, armed with the bytes of BOTH CTAs' loads
            mbar_init(empty + s, 1);     // the leader's multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full + b, 1);
            mbar_init(tmem_empty + b, 2 * kEpilogueWarps);   // the epilogue warps of BOTH CTAs
        }
        mbar_fence_init();
    }
    cluster_sync_all();                  // barrier words of the peer exist before anyone signals them
    if (warp == 1)
        tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t const tmem_base = *tmem_slot;
    cluster_sync_all();

    constexpr int kGroupM = 8;
    auto decode = [&](int item, int& m0, int& n0) {
        int const group_size = kGroupM * n_tiles;
        int const first_m = (item / group_size) * kGroupM;
        int const gm = min(m_tiles - first_m, kGroupM);
        int const in_group = item % group_size;
        m0 = (first_m + in_group % gm) * 2 * kBM;
        n0 = (in_group / gm) * BN;
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = pair; item < total_items; item += pairs) {
                int m0, n0;
                decode(item, m0, n0);
                for (int i = 0; i < total_kblocks; ++i, ++it) {
                    int const s = it % Cfg::kStages;
                    mbar_wait(empty + s, ((it / Cfg::kStages) & 1) ^ 1, 31);
                    uint32_t const leader_full = mapa_shared(smem_u32(full + s), 0);
                    uint8_t* a_dst = smem + s * Cfg::kStageBytes;
                    // Only the leader arms the barrier (a local arrive: a remote `arrive.release.cluster` per k-block costs the producer a
                    // MEMBAR + ERRBAR each time, 60 % of its cycles in the first version of this kernel). The peer's bytes may land
                    // before the leader's expect_tx: the transaction count goes negative, the phase still needs the leader's arrive.
                    if (leader)
                        mbar_expect_tx(full + s, 2 * Cfg::kStageBytes);
                    tma_load_2d_2sm(a_dst, &tmap_a, leader_full, i * kBK, m0 + static_cast<int>(rank) * kBM);
                    tma_load_2d_2sm(a_dst + Cfg::kABytes, &tmap_b, leader_full, i * kBK, n0 + static_cast<int>(rank) * (BN / 2));
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc(2 * kBM, BN, false, false, 1);
            uint32_t it = 0, j = 0;
            for (int item = pair; item < total_items; item += pairs, ++j) {
                uint32_t const buf = j & 1;
                mbar_wait(tmem_empty + buf, ((j >> 1) & 1) ^ 1, 32);
                tc_fence_after();
                uint32_t const acc = tmem_base + buf * BN;
                for (int i = 0; i < total_kblocks; ++i, ++it) {
                    int const s = it % Cfg::kStages;
                    mbar_wait(full + s, (it / Cfg::kStages) & 1, 33);
                    tc_fence_after();
                    uint32_t const a_addr = smem_u32(smem + s * Cfg::kStageBytes), b_addr = a_addr + Cfg::kABytes;
#pragma unroll
                    for (int kk = 0; kk < kBK / kUmmaK; ++kk)
                        umma_f16_2sm(acc, umma_smem_desc(a_addr + kk * 32, 16, 1024), umma_smem_desc(b_addr + kk * 32, 16, 1024), idesc, !(i == 0 && kk == 0));
                    umma_commit_2sm(empty + s, 0b11);
                }
                umma_commit_2sm(tmem_full + buf, 0b11);
            }
        }
    } else {
        uint32_t j = 0;
        for (int item = pair; item < total_items; item += pairs, ++j) {
            int m0, n0;
            decode(item, m0, n0);
            uint32_t const buf = j & 1;
            mbar_wait(tmem_full + buf, (j >> 1) & 1, 34);
            tc_fence_after();
            int const ewarp = (warp & 3) | (((warp - 2) >> 2) << 2);
            int const row = m0 + static_cast<int>(rank) * kBM + (warp & 3) * 32 + lane;
            epilogue_rows_staged<BN>(p, tmem_base + buf * BN, ewarp, lane, row < p.M, static_cast<long long>(row) * p.ldc, n0, epi_stage);
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive_cluster(mapa_shared(smem_u32(tmem_empty + buf), 0));
        }
    }
    tc_fence_before();
    cluster_sync_all();     // the peer may still be reading operands / signalling barriers of this CTA
    if (warp == 1)
        tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
}

template<int BN>
int launch_pair_gemm(CUtensorMap const& ta, CUtensorMap const& tb, GemmParams const& p, cudaStream_t stream) {
    using Cfg = PairConfig<BN>;
    auto kernel = gemm_tcgen05_pair_kernel<BN>;
    static bool configured = false;
    static int sms = 0;
    if (!configured) {
        AGB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        int device = 0;
        AGB_CUDA_OK(cudaGetDevice(&device));
        AGB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        configured = true;
    }
    int const m_tiles = (p.M + 2 * kBM - 1) / (2 * kBM), n_tiles = (p.N + BN - 1) / BN;
    long long const items = static_cast<long long>(m_tiles) * n_tiles;
    int const pairs = static_cast<int>(items < sms / 2 ? items : sms / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kPersistentThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AGB_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, ta, tb, p, m_tiles, n_tiles));
    return 0;
}

} // namespace
