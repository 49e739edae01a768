#pragma once
// bf16 GEMM on the 5th-generation tensor cores: C[M,N] = op(A) * op(B) with fp32 accumulation in TMEM.
//
// One kernel family covers the three products a layer needs:
//   forward   Y  = X  * W^T      A K-major   ([M,K] rows)      B K-major  ([N,K] rows)
//   dgrad     dX = dY * W        A K-major   ([M,K] rows)      B MN-major ([K,N] rows, N contiguous)
//   wgrad     dW = dY^T * X      A MN-major  ([K,M] rows)      B MN-major ([K,N] rows)   (+ split-K)
// so activations and weights are consumed in the layout they already have: no transposes.
//
// Structure (one 128 x BN output tile per CTA, two CTAs per SM so that one CTA's epilogue overlaps the
// other's main loop):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B swizzle) into a kStages-deep smem ring,
//               completion signalled on `full` mbarriers (complete_tx::bytes)
//   warp 1      allocates TMEM, one elected lane issues tcgen05.mma (kind::f16, M=128, N=BN, K=16) from the
//               smem descriptors, tcgen05.commit releases ring slots (`empty`) and finally arms `tmem_full`
//   warps 2..5  epilogue: tcgen05.ld 32x32b -> registers, + bias, ReLU, convert, store (or red.add for split-K)
//
// blockIdx.z = K split; each split handles a contiguous range of 64-wide K blocks and, when splits > 1,
// accumulates with fp32 atomics into a pre-zeroed C.

#include <cuda_bf16.h>
#include <cstdlib>

#include <agb_device.cuh>
#include <agb_tcgen05.cuh>

using namespace agb;
using namespace agb::sm100;

namespace {

constexpr int kBM = 128;       // UMMA M
constexpr int kBK = 64;        // bf16: K elements per stage = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;

// Operand element types. A pipeline stage is always 128 bytes of K per row (one swizzle row) and one tcgen05.mma consumes 32 bytes of it:
//   bf16  kind::f16   64 K-elements per stage, 16 per instruction, MN-major chunks of 64 elements
//   tf32  kind::tf32  32 K-elements per stage,  8 per instruction, MN-major chunks of 32 elements - fp32 operands straight from memory
//         (the tensor core reads the top 19 bits): the parity-precision path for the fp32 reference (`graph.py:267-273`)
struct ElemBF16 {
    using type = __nv_bfloat16;
    static constexpr int kBytes = 2, kBK = 64, kUmmaK = 16, kChunk = 64;
    static constexpr uint32_t kFormat = 1;   // instruction-descriptor A/B format: BF16
    static constexpr bool kTf32 = false;
};
struct ElemTF32 {
    using type = float;
    static constexpr int kBytes = 4, kBK = 32, kUmmaK = 8, kChunk = 32;
    static constexpr uint32_t kFormat = 2;   // TF32
    static constexpr bool kTf32 = true;
};

template<int BN> struct Config {
    static constexpr int kStages = BN <= 64 ? 4 : 3;   // <= 97 KB of smem for BN <= 128: two CTAs per SM
    static constexpr uint32_t kABytes = kBM * 128;   // 128 rows x 128 bytes of K (either element type)
    static constexpr uint32_t kBBytes = BN * 128;
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
    static constexpr uint32_t kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
};

struct GemmParams {
    int M, N, K;            // logical problem
    long long ldc;          // row stride of C (elements)
    void* C;
    float const* bias;      // per-N fp32 bias or null
    int relu;
    int out_fp32;           // C element type: 1 = float, 0 = bf16
    int atomic;             // accumulate with red.add.f32 (split-K); requires out_fp32
    int kblocks_per_split;
    int groups;              // grouped weight gradients: independent K ranges (one per logical worker) ...
    long long c_group_stride; // ... each accumulating into C + group * c_group_stride
};

// Epilogue of one 128 x BN accumulator tile: TMEM -> registers -> (+bias, ReLU, convert) -> global memory.
// `row_valid` / `row_offset`: whether this thread's accumulator row (TMEM lane quarter * 32 + lane) is a real output row
// and the element offset of that row in C (lets implicit-GEMM convolutions scatter pixel-box rows to NHWC addresses).
template<int BN>
__device__ __forceinline__ void epilogue_rows(GemmParams const& p, uint32_t tmem_acc, int warp, bool row_valid, long long row_offset, int n0) {
    int const quarter = warp & 3;                 // TMEM lane quarter this warp may access
    uint32_t const taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
        float v[32];
        tmem_ld_32x32(taddr + c, v);
        int const col0 = n0 + c;
        if (row_valid && col0 < p.N) {
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < p.N)
                        v[j] += __ldg(p.bias + col0 + j);
            }
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = fmaxf(v[j], 0.f);
            }
            bool const full_span = col0 + 32 <= p.N;
            if (p.out_fp32) {
                float* dst = static_cast<float*>(p.C) + row_offset + col0;
                if (p.atomic) {
                    if (full_span && (p.ldc & 3) == 0) {   // 128-bit fp32 reductions (sm_90+)
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.N)
                                atomicAdd(dst + j, v[j]);
                    }
                } else if (full_span && (p.ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < p.N)
                            dst[j] = v[j];
                }
            } else {
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.C) + row_offset + col0;
                if (full_span && (p.ldc & 7) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                        uint4 packed = make_uint4(*reinterpret_cast<unsigned*>(&h0), *reinterpret_cast<unsigned*>(&h1), *reinterpret_cast<unsigned*>(&h2), *reinterpret_cast<unsigned*>(&h3));
                        *reinterpret_cast<uint4*>(dst + j) = packed;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < p.N)
                            dst[j] = __float2bfloat16(v[j]);
                }
            }
        }
    }
}

// Coalesced variant of the epilogue, run by 8 warps: two warps share each TMEM lane quarter and split the tile's columns.
// Every warp owns a 32-row x 128-byte staging tile in shared memory: accumulator rows (thread = row) are loaded from TMEM
// (all loads of a pass in flight before one wait), converted and written to the tile, then read back so that consecutive
// lanes cover contiguous bytes of a row: each global store (or fp32 vector reduction) writes full 64/128-byte row
// segments instead of 32 scattered 16-byte pieces. Falls back to the per-thread path when N / ldc forbid 16-byte pieces.
constexpr int kStagePitch = 144;                      // 128 B of payload + 16 B: conflict-free for both access patterns
constexpr int kStageBytesPerWarp = 32 * kStagePitch;
constexpr int kEpilogueWarps = 8;
constexpr int kStageBytes = kEpilogueWarps * kStageBytesPerWarp;
constexpr int kPersistentThreads = 64 + kEpilogueWarps * 32;   // TMA warp + MMA warp + 8 epilogue warps

// `ewarp` in [0, 8): quarter = ewarp & 3 (must equal the hardware warp id & 3), column half = ewarp >> 2.
template<int BN>
__device__ __forceinline__ void epilogue_rows_staged(GemmParams const& p, uint32_t tmem_acc, int ewarp, int lane, bool row_valid, long long row_offset, int n0, uint8_t* stage_base) {
    int const quarter = ewarp & 3, half = ewarp >> 2;
    constexpr int kHalfCols = BN / 2;
    bool const fast = (p.out_fp32 ? ((p.ldc & 3) == 0 && (p.N & 3) == 0) : ((p.ldc & 7) == 0 && (p.N & 7) == 0)) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    uint32_t const taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + half * kHalfCols;
    int const nbase = n0 + half * kHalfCols;
    if (!fast) {   // per-thread stores, 32 columns at a time
#pragma unroll 1
        for (int c = 0; c < kHalfCols; c += 32) {
            float v[32];
            tmem_ld_32x32(taddr + c, v);
            int const col0 = nbase + c;
            if (row_valid && col0 < p.N) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (col0 + j < p.N) {
                        float val = v[j] + (p.bias ? __ldg(p.bias + col0 + j) : 0.f);
                        if (p.relu)
                            val = fmaxf(val, 0.f);
                        if (p.out_fp32) {
                            float* dst = static_cast<float*>(p.C) + row_offset + col0 + j;
                            if (p.atomic)
                                atomicAdd(dst, val);
                            else
                                *dst = val;
                        } else {
                            static_cast<__nv_bfloat16*>(p.C)[row_offset + col0 + j] = __float2bfloat16(val);
                        }
                    }
                }
            }
        }
        return;
    }
    uint8_t* stage = stage_base + ewarp * kStageBytesPerWarp;
    unsigned const off_lo = static_cast<unsigned>(row_offset), off_hi = static_cast<unsigned>(static_cast<unsigned long long>(row_offset) >> 32);
    if (p.out_fp32) {
        // 32 fp32 columns (128 B per row) per pass
#pragma unroll 1
        for (int c = 0; c < kHalfCols; c += 32) {
            int const col0 = nbase + c;
            if (col0 >= p.N)
                break;
            float v[32];
            tmem_ld_32x32(taddr + c, v);
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < p.N)
                        v[j] += __ldg(p.bias + col0 + j);
            }
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = fmaxf(v[j], 0.f);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(stage + lane * kStagePitch + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
            int const piece = lane & 7, sub = lane >> 3;
            int const piece_col = col0 + piece * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int const r = 4 * i + sub;
                bool const valid = __shfl_sync(0xffffffffu, static_cast<int>(row_valid), r) != 0;
                unsigned const lo = __shfl_sync(0xffffffffu, off_lo, r), hi = __shfl_sync(0xffffffffu, off_hi, r);
                long long const offset = static_cast<long long>((static_cast<unsigned long long>(hi) << 32) | lo);
                if (valid && piece_col < p.N) {
                    float4 const data = *reinterpret_cast<float4 const*>(stage + r * kStagePitch + piece * 16);
                    float* g = static_cast<float*>(p.C) + offset + piece_col;
                    if (p.atomic)
                        atomicAdd(reinterpret_cast<float4*>(g), data);
                    else
                        *reinterpret_cast<float4*>(g) = data;
                }
            }
            __syncwarp();
        }
    } else {
        // bf16: kPass columns per pass (64 when the half is wide enough: 128 B per row), both TMEM loads in flight together
        constexpr int kPass = kHalfCols >= 64 ? 64 : 32;
        constexpr int kPieces = kPass * 2 / 16;          // 16-byte pieces per row: 8 or 4
        constexpr int kRowsPerInstr = 32 / kPieces;
#pragma unroll 1
        for (int c = 0; c < kHalfCols; c += kPass) {
            int const col0 = nbase + c;
            if (col0 >= p.N)
                break;
            uint32_t raw[kPass];
            tmem_ld_32x32_async(taddr + c, *reinterpret_cast<uint32_t (*)[32]>(&raw[0]));
            if (kPass == 64)
                tmem_ld_32x32_async(taddr + c + 32, *reinterpret_cast<uint32_t (*)[32]>(&raw[kPass == 64 ? 32 : 0]));
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < kPass; j += 8) {
                float v[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    v[t] = __uint_as_float(raw[j + t]);
                    if (p.bias && col0 + j + t < p.N)
                        v[t] += __ldg(p.bias + col0 + j + t);
                    if (p.relu)
                        v[t] = fmaxf(v[t], 0.f);
                }
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
                *reinterpret_cast<uint4*>(stage + lane * kStagePitch + j * 2) = make_uint4(*reinterpret_cast<unsigned*>(&h0), *reinterpret_cast<unsigned*>(&h1), *reinterpret_cast<unsigned*>(&h2), *reinterpret_cast<unsigned*>(&h3));
            }
            __syncwarp();
            int const piece = lane % kPieces, sub = lane / kPieces;
            int const piece_col = col0 + piece * 8;
#pragma unroll
            for (int i = 0; i < kPieces; ++i) {
                int const r = kRowsPerInstr * i + sub;
                bool const valid = __shfl_sync(0xffffffffu, static_cast<int>(row_valid), r) != 0;
                unsigned const lo = __shfl_sync(0xffffffffu, off_lo, r), hi = __shfl_sync(0xffffffffu, off_hi, r);
                long long const offset = static_cast<long long>((static_cast<unsigned long long>(hi) << 32) | lo);
                if (valid && piece_col < p.N)
                    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.C) + offset + piece_col) = *reinterpret_cast<uint4 const*>(stage + r * kStagePitch + piece * 16);
            }
            __syncwarp();
        }
    }
}

template<int BN>
__device__ __forceinline__ void epilogue_tile(GemmParams const& p, uint32_t tmem_acc, int warp, int lane, int m0, int n0) {
    int const row = m0 + (warp & 3) * 32 + lane;
    epilogue_rows<BN>(p, tmem_acc, warp, row < p.M, static_cast<long long>(row) * p.ldc, n0);
}

// Loads of one pipeline stage (A and B tiles of the k-block starting at element `k`).
template<int BN, bool A_MN, bool B_MN, typename E = ElemBF16>
__device__ __forceinline__ void produce_stage(CUtensorMap const* tmap_a, CUtensorMap const* tmap_b, uint8_t* a_dst, uint64_t* bar, int m0, int n0, int k, int group = -1) {
    uint8_t* b_dst = a_dst + Config<BN>::kABytes;
    constexpr int kChunkBytes = E::kBK * 128;   // one MN-major chunk: kBK k-rows of 128 bytes
    mbar_expect_tx(bar, Config<BN>::kStageBytes);
    if (A_MN && B_MN && group >= 0) {   // grouped weight gradient: 3-D maps (inner, row in group, group)
#pragma unroll
        for (int c = 0; c < kBM / E::kChunk; ++c)
            tma_load_3d(a_dst + c * kChunkBytes, tmap_a, bar, m0 + c * E::kChunk, k, group);
#pragma unroll
        for (int c = 0; c < BN / E::kChunk; ++c)
            tma_load_3d(b_dst + c * kChunkBytes, tmap_b, bar, n0 + c * E::kChunk, k, group);
        return;
    }
    if (A_MN) { // rows = K index, kChunk M-elements per row; one box per chunk
#pragma unroll
        for (int c = 0; c < kBM / E::kChunk; ++c)
            tma_load_2d(a_dst + c * kChunkBytes, tmap_a, bar, m0 + c * E::kChunk, k);
    } else {
        tma_load_2d(a_dst, tmap_a, bar, k, m0);
    }
    if (B_MN) {
#pragma unroll
        for (int c = 0; c < BN / E::kChunk; ++c)
            tma_load_2d(b_dst + c * kChunkBytes, tmap_b, bar, n0 + c * E::kChunk, k);
    } else {
        tma_load_2d(b_dst, tmap_b, bar, k, n0);
    }
}

// tcgen05.mma over one staged k-block (4 instructions of 32 bytes of K each).
template<int BN, bool A_MN, bool B_MN, typename E = ElemBF16>
__device__ __forceinline__ void consume_stage(uint32_t a_addr, uint32_t tmem_acc, bool first) {
    constexpr uint32_t idesc = umma_idesc(kBM, BN, A_MN, B_MN, E::kFormat);
    uint32_t const b_addr = a_addr + Config<BN>::kABytes;
    constexpr int kChunkBytes = E::kBK * 128;
#pragma unroll
    for (int kk = 0; kk < E::kBK / E::kUmmaK; ++kk) {
        // K-major: one instruction's K = 32 B further inside the 128B swizzled row.
        // MN-major: kUmmaK k-rows = kUmmaK * 128 B further; chunks of kChunk MN-elements are kChunkBytes apart.
        // MN-major tf32: 32-byte-chunk swizzle over 4-row atoms (SWIZZLE_128B_BASE32B), k-groups of 4 rows 512 B apart.
        constexpr uint32_t kMnSbo = E::kTf32 ? 512 : 1024, kMnLayout = E::kTf32 ? 1 : 2;
        uint64_t const da = A_MN ? umma_smem_desc(a_addr + kk * E::kUmmaK * 128, kChunkBytes, kMnSbo, kMnLayout) : umma_smem_desc(a_addr + kk * 32, 16, 1024);
        uint64_t const db = B_MN ? umma_smem_desc(b_addr + kk * E::kUmmaK * 128, kChunkBytes, kMnSbo, kMnLayout) : umma_smem_desc(b_addr + kk * 32, 16, 1024);
        if (E::kTf32)
            umma_tf32(tmem_acc, da, db, idesc, !(first && kk == 0));
        else
            umma_f16(tmem_acc, da, db, idesc, !(first && kk == 0));
    }
}

// ---------------------------------------------------------------------------------------------------------------- //
// Persistent variant: one CTA per SM loops over (tile, k-split) work items. The smem ring runs continuously across
// items, and the accumulator is double-buffered in TMEM (2 x BN columns) so that the epilogue of item j overlaps the
// main loop of item j + 1; barrier setup and the TMEM allocation are paid once per CTA instead of once per tile.
template<int BN> struct PersistentConfig {
    static constexpr int kStages = BN <= 64 ? 7 : BN <= 128 ? 5 : 3;   // ring + 36 KB of epilogue staging <= 227 KB
    static constexpr uint32_t kSmemBytes = kStages * Config<BN>::kStageBytes + 1024 + 256 + kStageBytes;   // + epilogue staging tiles
    static constexpr uint32_t kTmemCols = 2 * Config<BN>::kTmemCols;   // <= 512
};

template<int BN, bool A_MN, bool B_MN, typename E = ElemBF16>
__global__ void __launch_bounds__(kPersistentThreads, 1) gemm_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, GemmParams const p, int m_tiles, int n_tiles, int splits) {
    using Cfg = Config<BN>;
    using PCfg = PersistentConfig<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + PCfg::kStages * Cfg::kStageBytes);
    uint64_t* empty = full + PCfg::kStages;
    uint64_t* tmem_full = empty + PCfg::kStages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint8_t* epi_stage = smem + PCfg::kStages * Cfg::kStageBytes + 256;

    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const total_kblocks = (p.K + E::kBK - 1) / E::kBK;
    int const total_items = m_tiles * n_tiles * splits * (p.groups > 1 ? p.groups : 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < PCfg::kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full + b, 1);
            mbar_init(tmem_empty + b, kEpilogueWarps);   // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1)
        tmem_alloc<PCfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t const tmem_base = *tmem_slot;
    pdl_trigger();   // set-up done: the next kernel may start its own; ...
    pdl_wait();      // ... our inputs are complete only once the previous kernel has finished

    // item -> tile in grouped order (panels of 16 m-tiles, m fastest inside a panel, then n): the ~148 tiles in flight
    // cover a ~16 x 9 patch, so every A and B tile fetched from L2/HBM is reused by several CTAs of the same wave.
    constexpr int kGroupM = 16;
    int const items_per_group = m_tiles * n_tiles * splits;
    auto decode = [&](int item, int& m0, int& n0, int& kb_begin, int& nkb, int& group) {
        group = item / items_per_group;
        item -= group * items_per_group;
        int const tile = item % (m_tiles * n_tiles), split = item / (m_tiles * n_tiles);
        int const group_size = kGroupM * n_tiles;
        int const first_m = (tile / group_size) * kGroupM;
        int const gm = min(m_tiles - first_m, kGroupM);
        int const in_group = tile % group_size;
        m0 = (first_m + in_group % gm) * kBM;
        n0 = (in_group / gm) * BN;
        kb_begin = split * p.kblocks_per_split;
        int const kb_end = min(total_kblocks, kb_begin + p.kblocks_per_split);
        nkb = kb_end - kb_begin;
    };
    bool const grouped = p.groups > 1 || (A_MN && B_MN);   // TN products always use the 3-D maps

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;   // running k-block counter => ring slot and phase
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                int m0, n0, kb_begin, nkb, group;
                decode(item, m0, n0, kb_begin, nkb, group);
                for (int i = 0; i < nkb; ++i, ++it) {
                    int const s = it % PCfg::kStages;
                    mbar_wait(empty + s, ((it / PCfg::kStages) & 1) ^ 1, 11);
                    produce_stage<BN, A_MN, B_MN, E>(&tmap_a, &tmap_b, smem + s * Cfg::kStageBytes, full + s, m0, n0, (kb_begin + i) * E::kBK, grouped ? group : -1);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, j = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++j) {
                int m0, n0, kb_begin, nkb, group;
                decode(item, m0, n0, kb_begin, nkb, group);
                uint32_t const buf = j & 1;
                mbar_wait(tmem_empty + buf, ((j >> 1) & 1) ^ 1, 12);   // epilogue drained this accumulator
                tc_fence_after();
                uint32_t const acc = tmem_base + buf * Cfg::kTmemCols;
                for (int i = 0; i < nkb; ++i, ++it) {
                    int const s = it % PCfg::kStages;
                    mbar_wait(full + s, (it / PCfg::kStages) & 1, 13);
                    tc_fence_after();
                    consume_stage<BN, A_MN, B_MN, E>(smem_u32(smem + s * Cfg::kStageBytes), acc, i == 0);
                    umma_commit(empty + s);
                }
                umma_commit(tmem_full + buf);
            }
        }
    } else {
        uint32_t j = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++j) {
            int m0, n0, kb_begin, nkb, group;
            decode(item, m0, n0, kb_begin, nkb, group);
            uint32_t const buf = j & 1;
            mbar_wait(tmem_full + buf, (j >> 1) & 1, 14);
            tc_fence_after();
            {
                int const ewarp = ((warp & 3)) | (((warp - 2) >> 2) << 2);   // quarter from the hardware warp id, half from the warp's group
                int const row = m0 + (warp & 3) * 32 + lane;
                epilogue_rows_staged<BN>(p, tmem_base + buf * Cfg::kTmemCols, ewarp, lane, row < p.M, static_cast<long long>(row) * p.ldc + group * p.c_group_stride, n0, epi_stage);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(tmem_empty + buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        tmem_dealloc<PCfg::kTmemCols>(tmem_base);
}

template<int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, BN <= 128 ? 2 : 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, GemmParams const p) {
    using Cfg = Config<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty = full + Cfg::kStages;
    uint64_t* tmem_full = empty + Cfg::kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
    int const total_kblocks = (p.K + kBK - 1) / kBK;
    int const kb_begin = blockIdx.z * p.kblocks_per_split;
    int const kb_end = min(total_kblocks, kb_begin + p.kblocks_per_split);
    int const nkb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1)
        tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t const tmem_base = *tmem_slot;
    pdl_trigger();   // set-up done: the next kernel may start its own; ...
    pdl_wait();      // ... our inputs are complete only once the previous kernel has finished

    if (nkb > 0) {
        if (warp == 0) {
            // ------------------------------ TMA producer ------------------------------ //
            if (lane == 0) {
                for (int i = 0; i < nkb; ++i) {
                    int const s = i % Cfg::kStages;
                    uint32_t const phase = (i / Cfg::kStages) & 1;
                    mbar_wait(empty + s, phase ^ 1, 1);
                    produce_stage<BN, A_MN, B_MN>(&tmap_a, &tmap_b, smem + s * Cfg::kStageBytes, full + s, m0, n0, (kb_begin + i) * kBK, (A_MN && B_MN) ? 0 : -1);
                }
            }
        } else if (warp == 1) {
            // ------------------------------ MMA issuer ------------------------------ //
            if (lane == 0) {
                for (int i = 0; i < nkb; ++i) {
                    int const s = i % Cfg::kStages;
                    uint32_t const phase = (i / Cfg::kStages) & 1;
                    mbar_wait(full + s, phase, 2);
                    tc_fence_after();
                    consume_stage<BN, A_MN, B_MN>(smem_u32(smem + s * Cfg::kStageBytes), tmem_base, i == 0);
                    umma_commit(empty + s);      // slot reusable once these MMAs have read it
                }
                umma_commit(tmem_full);          // accumulator complete
            }
        } else {
            // ------------------------------ epilogue ------------------------------ //
            mbar_wait(tmem_full, 0, 3);
            tc_fence_after();
            epilogue_tile<BN>(p, tmem_base, warp, lane, m0, n0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

static int g_persistent = -1;   // -1: read AGB_GEMM_PERSISTENT at first use

template<int BN, bool A_MN, bool B_MN, typename E = ElemBF16>
int launch_gemm(CUtensorMap const& ta, CUtensorMap const& tb, GemmParams const& p, int splits, cudaStream_t stream) {
    using Cfg = Config<BN>;
    if (g_persistent < 0) {
        char const* env = std::getenv("AGB_GEMM_PERSISTENT");
        g_persistent = env ? std::atoi(env) : 1;
    }
    int const m_tiles = (p.M + kBM - 1) / kBM, n_tiles = (p.N + BN - 1) / BN;
    if (g_persistent || E::kTf32) {   // the TF32 path only exists as the persistent kernel
        using PCfg = PersistentConfig<BN>;
        auto kernel = gemm_tcgen05_persistent_kernel<BN, A_MN, B_MN, E>;
        static bool configured = false;
        static int sms = 0;
        if (!configured) {
            AGB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg::kSmemBytes));
            int device = 0;
            AGB_CUDA_OK(cudaGetDevice(&device));
            AGB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            configured = true;
        }
        long long const items = static_cast<long long>(m_tiles) * n_tiles * splits * (p.groups > 1 ? p.groups : 1);
        int const grid = static_cast<int>(items < sms ? items : sms);
        AGB_CUDA_OK(launch_pdl(kernel, dim3(grid), dim3(kPersistentThreads), PCfg::kSmemBytes, stream, ta, tb, p, m_tiles, n_tiles, splits));
        return 0;
    }
    if (p.groups > 1 || E::kTf32)
        return 207;   // grouped / TF32 products need the persistent kernel
    auto kernel = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
    static bool configured = false;
    if (!configured) {
        AGB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        configured = true;
    }
    dim3 grid(m_tiles, n_tiles, splits);
    kernel<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
    AGB_CUDA_OK(cudaGetLastError());
    return 0;
}

template<bool A_MN, bool B_MN>
int dispatch_bn(int bn, CUtensorMap const& ta, CUtensorMap const& tb, GemmParams const& p, int splits, cudaStream_t stream) {
    switch (bn) {
        case 64: return launch_gemm<64, A_MN, B_MN>(ta, tb, p, splits, stream);
        case 128: return launch_gemm<128, A_MN, B_MN>(ta, tb, p, splits, stream);
        case 256: return launch_gemm<256, A_MN, B_MN>(ta, tb, p, splits, stream);
    }
    return 203;
}

template<bool A_MN, bool B_MN>
int dispatch_bn_tf32(int bn, CUtensorMap const& ta, CUtensorMap const& tb, GemmParams const& p, int splits, cudaStream_t stream) {
    switch (bn) {
        case 64: return launch_gemm<64, A_MN, B_MN, ElemTF32>(ta, tb, p, splits, stream);
        case 128: return launch_gemm<128, A_MN, B_MN, ElemTF32>(ta, tb, p, splits, stream);
    }
    return 203;
}

} // namespace

