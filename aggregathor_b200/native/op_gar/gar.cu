// Fused gather + gradient-aggregation-rule + optimizer update + parameter broadcast (sm_100a).
//
// This is the B200 replacement of the reference's whole PS data path
// (worker->PS gradient transfer, `native/op_krum/cpu.cpp`, `native/op_bulyan/cpu.cpp`,
// `deprecated_native/native.cpp` median/averaged-median/average-nan, `opt.apply_gradients`,
// PS->worker variable transfer). One persistent cooperative kernel per rank:
//
//   entry barrier   all ranks have published their workers' gradients (system-scope flags)
//   phase A         (Krum/Bulyan) stream the owned coordinate slice of all n gradients straight
//                   from the peers' buffers (P2P loads over NVLink), accumulate the n(n-1)/2
//                   partial squared distances with direct differences, stage the tile locally
//   exchange        per-rank partial matrices go to every peer's mailbox; summed in rank order
//                   => bit-identical distance matrix on every rank
//   select          one warp: Krum scores / Bulyan iterative selection (replicated on all ranks)
//   phase D         aggregate the slice (mean of selected / coordinate-wise trimmed mean / median /
//                   NaN-aware mean), apply the optimizer on the slice, store the new parameters
//                   into every rank's parameter buffer (P2P stores or one NVLS multimem.st)
//   exit barrier    every rank's slice has landed everywhere; gradient buffers may be reused
//
// With R = 1 the same kernel is the stand-alone `[n, d] -> [d]` aggregation op.
// Ordering convention: finite ascending, non-finite last, ties -> lower worker index.

#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include <agb_device.cuh>

namespace cg = cooperative_groups;
using namespace agb;

namespace {

constexpr int kMaxWorkers = 16;  // register-resident rules
constexpr int kMaxRanks = 16;
constexpr int kMaxPairs = kMaxWorkers * (kMaxWorkers - 1) / 2;

enum Rule { kAverage = 0, kAverageNan = 1, kMedian = 2, kAveragedMedian = 3, kKrum = 4, kBulyan = 5 };
enum Opt { kNone = 0, kSgd = 1, kAdam = 2, kRmsprop = 3, kAdagrad = 4, kAdadelta = 5 };

struct GarArgs {
    int n, f, m, beta, rule;
    int R, rank;
    long long lo, hi;                  // owned coordinate slice, multiples of 4
    float const* grad[kMaxWorkers];    // row base pointers (local or peer-mapped)
    float const* grad_mc;              // multicast address of the [w, d] gradient matrix (NVLS in-switch reduction), or null
    int workers_per_rank;
    long long row_stride;              // elements between two rows of a rank's gradient matrix
    float* agg_out;                    // optional [d] (local): aggregated gradient of the slice
    int opt;
    float lr, h0, h1, h2;              // adam: b1,b2,eps | rmsprop: decay,momentum,eps | adadelta: rho,eps
    float* param;                      // local fp32 parameters [d]
    float* slot0;
    float* slot1;
    float* param_dst[kMaxRanks];       // every rank's parameter buffer (peer-mapped), incl. own
    float* param_mc;                   // multicast address of the parameter buffers, or null
    __nv_bfloat16* param_bf16_dst[kMaxRanks]; // optional bf16 compute copy of the parameters
    uint32_t* signal[kMaxRanks];       // [3][R] flags of every rank
    float* mailbox[kMaxRanks];         // [R][kMaxPairs] partial distances of every rank
    uint32_t epoch;
    float* cta_partials;               // [grid][kMaxPairs]
    float* staging;                    // [n][hi - lo] or null
    float* dist_out;                   // optional [n * n]
    int* info;                         // optional [64]: selection masks for tests/diagnostics
};

struct Shared {
    float dist[kMaxWorkers][kMaxWorkers + 1];
    float pruned[kMaxWorkers][kMaxWorkers + 1];
    float scores[kMaxWorkers];
    float warp_partials[16][kMaxPairs];
    unsigned selmask[kMaxWorkers];   // Krum: [0]; Bulyan: one per round
    int selcount[kMaxWorkers];
    int theta;
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float& f4_at(float4& v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; }

// ---- optimizer update of 4 consecutive coordinates + broadcast ------------- //
__device__ __forceinline__ void apply_update(GarArgs const& a, long long x, float4 g) {
    if (a.agg_out)
        *reinterpret_cast<float4*>(a.agg_out + x) = g;
    if (a.opt == kNone)
        return;
    float4 p = *reinterpret_cast<float4 const*>(a.param + x);
    if (a.opt == kSgd) {
        p.x -= a.lr * g.x; p.y -= a.lr * g.y; p.z -= a.lr * g.z; p.w -= a.lr * g.w;
    } else {
        float4 s0 = *reinterpret_cast<float4 const*>(a.slot0 + x);
        float4 s1 = a.slot1 ? *reinterpret_cast<float4 const*>(a.slot1 + x) : f4_zero();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float gc = f4_at(g, c), &pc = f4_at(p, c), &u = f4_at(s0, c), &v = f4_at(s1, c);
            if (a.opt == kAdam) {            // lr already carries the bias correction sqrt(1-b2^t)/(1-b1^t)
                u = a.h0 * u + (1.f - a.h0) * gc;
                v = a.h1 * v + (1.f - a.h1) * gc * gc;
                pc -= a.lr * u / (sqrtf(v) + a.h2);
            } else if (a.opt == kRmsprop) {  // u: mean square, v: momentum
                u = a.h0 * u + (1.f - a.h0) * gc * gc;
                v = a.h1 * v + a.lr * gc * rsqrtf(u + a.h2);
                pc -= v;
            } else if (a.opt == kAdagrad) {  // u: accumulator
                u += gc * gc;
                pc -= a.lr * gc * rsqrtf(u);
            } else {                         // adadelta; u: accum, v: accum_update
                u = a.h0 * u + (1.f - a.h0) * gc * gc;
                float upd = sqrtf(v + a.h1) * rsqrtf(u + a.h1) * gc;
                v = a.h0 * v + (1.f - a.h0) * upd * upd;
                pc -= a.lr * upd;
            }
        }
        *reinterpret_cast<float4*>(a.slot0 + x) = s0;
        if (a.slot1)
            *reinterpret_cast<float4*>(a.slot1 + x) = s1;
    }
    if (a.param_mc) {
        multimem_st_f4(a.param_mc + x, p);
    } else {
        for (int q = 0; q < a.R; ++q)
            st_stream_f4(a.param_dst[q] + x, p);
    }
    if (a.param_bf16_dst[0]) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
        uint2 packed = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
        for (int q = 0; q < a.R; ++q)
            *reinterpret_cast<uint2*>(a.param_bf16_dst[q] + x) = packed;
    }
}

// ---- coordinate-wise rules on N register-resident values -------------------- //
// rank-counting selection: branch free, NaN-correct, index-stable; O(n^2) compares per coordinate.
template<int N> __device__ __forceinline__ float coord_median(float const (&v)[N], int n) {
    float out = 0.f;
    int const target = n / 2;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i < n) {
            int rank = 0;
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (j < n && j != i)
                    rank += before(v[j], j, v[i], i) ? 1 : 0;
            if (rank == target)
                out = v[i];
        }
    }
    return out;
}

// mean of the `beta` values closest to the (upper) median, summed in index order
template<int N> __device__ __forceinline__ float coord_averaged_median(float const (&v)[N], int n, int beta) {
    float const zero = coord_median<N>(v, n);
    float dev[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
        dev[i] = fabsf(v[i] - zero);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i < n) {
            int rank = 0;
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (j < n && j != i)
                    rank += before(dev[j], j, dev[i], i) ? 1 : 0;
            if (rank < beta)
                sum += v[i];
        }
    }
    return sum / static_cast<float>(beta);
}

template<int N> __device__ __forceinline__ float coord_average_nan(float const (&v)[N], int n) {
    float sum = 0.f;
    int count = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (i < n && is_finite(v[i])) {
            sum += v[i];
            ++count;
        }
    return sum / static_cast<float>(count);
}

// ---- cross-rank flag barrier ---------------------------------------------------- //
__device__ __forceinline__ void signal_all(GarArgs const& a, int slot) {
    if (threadIdx.x < a.R)
        st_release_sys(a.signal[threadIdx.x] + slot * a.R + a.rank, a.epoch);
}
__device__ __forceinline__ void wait_all(GarArgs const& a, int slot) {
    if (threadIdx.x < a.R)
        wait_flag_sys(a.signal[a.rank] + slot * a.R + threadIdx.x, a.epoch);
}

// ---- selection stages (warp 0) -------------------------------------------------- //
__device__ void select_krum(GarArgs const& a, Shared& sh) {
    int const lane = threadIdx.x, n = a.n, count = a.n - a.f - 2;
    float score = 0.f;
    if (lane < n) {
        for (int r = 0; r < count; ++r) {  // add the distances in ascending order (matches the host oracle)
            for (int j = 0; j < n; ++j) {
                if (j == lane)
                    continue;
                int rank = 0;
                for (int k = 0; k < n; ++k)
                    if (k != lane && k != j)
                        rank += before(sh.dist[lane][k], k, sh.dist[lane][j], j) ? 1 : 0;
                if (rank == r)
                    score += sh.dist[lane][j];
            }
        }
        sh.scores[lane] = score;
    }
    __syncwarp();
    bool selected = false;
    if (lane < n) {
        int rank = 0;
        for (int j = 0; j < n; ++j)
            if (j != lane)
                rank += before(sh.scores[j], j, score, lane) ? 1 : 0;
        selected = rank < a.m;
    }
    unsigned mask = __ballot_sync(0xffffffffu, selected);
    if (lane == 0) {
        sh.selmask[0] = mask;
        sh.selcount[0] = a.m;
        sh.theta = 1;
    }
}

__device__ void select_bulyan(GarArgs const& a, Shared& sh) {
    int const lane = threadIdx.x, n = a.n, inscore = a.n - a.f - 2, theta = a.n - 2 * a.f - 2;
    float score = 0.f;
    if (lane < n) {
        for (int j = 0; j < n; ++j) {
            if (j == lane) {
                sh.pruned[lane][j] = 0.f;
                continue;
            }
            int rank = 0;
            for (int k = 0; k < n; ++k)
                if (k != lane && k != j)
                    rank += before(sh.dist[lane][k], k, sh.dist[lane][j], j) ? 1 : 0;
            sh.pruned[lane][j] = rank < inscore ? sh.dist[lane][j] : 0.f; // farthest f+1 never counted => never subtracted
        }
        for (int r = 0; r < inscore; ++r)
            for (int j = 0; j < n; ++j) {
                if (j == lane)
                    continue;
                int rank = 0;
                for (int k = 0; k < n; ++k)
                    if (k != lane && k != j)
                        rank += before(sh.dist[lane][k], k, sh.dist[lane][j], j) ? 1 : 0;
                if (rank == r)
                    score += sh.dist[lane][j];
            }
    }
    bool removed = false;
    for (int k = 0; k < theta; ++k) {
        if (lane < n)
            sh.scores[lane] = score;
        unsigned removed_mask = __ballot_sync(0xffffffffu, removed);
        __syncwarp();
        int rank = 0;
        if (lane < n) {
            for (int j = 0; j < n; ++j) {
                if (j == lane)
                    continue;
                bool rj = (removed_mask >> j) & 1u;
                bool j_first = rj != removed ? !rj : before(sh.scores[j], j, score, lane);
                rank += j_first ? 1 : 0;
            }
        }
        int const count = a.m - k;
        unsigned mask = __ballot_sync(0xffffffffu, lane < n && rank < count);
        unsigned best_mask = __ballot_sync(0xffffffffu, lane < n && rank == 0);
        int const best = __ffs(best_mask) - 1;
        if (lane == 0) {
            sh.selmask[k] = mask;
            sh.selcount[k] = count;
        }
        if (lane == best)
            removed = true;
        else if (lane < n && !removed)
            score -= sh.pruned[lane][best];
        __syncwarp();
    }
    if (lane == 0)
        sh.theta = theta;
}

// ---- the kernel ------------------------------------------------------------------ //
template<int N> __global__ void __launch_bounds__(N <= 8 ? 512 : 256, 1) gar_fused_kernel(GarArgs const a) {
    constexpr int NP = N * (N - 1) / 2;
    __shared__ Shared sh;
    cg::grid_group grid = cg::this_grid();
    int const n = a.n;
    long long const len = a.hi - a.lo, len4 = len >> 2;
    long long const tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    long long const nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    bool const multi = a.R > 1;
    bool const distance_rule = a.rule == kKrum || a.rule == kBulyan;

    if (multi) { // entry barrier: every rank's gradients are published
        if (blockIdx.x == 0)
            signal_all(a, 0);
        wait_all(a, 0);
        __syncthreads();
    }

    if (distance_rule) {
        // -------- phase A: partial pairwise squared distances over the owned slice -------- //
        float acc[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p)
            acc[p] = 0.f;
        for (long long v = tid; v < len4; v += nthreads) {
            long long const x = a.lo + (v << 2);
            float4 g[N];
#pragma unroll
            for (int i = 0; i < N; ++i)
                g[i] = i < n ? ld_stream_f4(a.grad[i] + x) : f4_zero();
            if (a.staging) {
#pragma unroll
                for (int i = 0; i < N; ++i)
                    if (i < n)
                        *reinterpret_cast<float4*>(a.staging + i * len + (v << 2)) = g[i];
            }
            int p = 0;
#pragma unroll
            for (int i = 0; i < N - 1; ++i) {
#pragma unroll
                for (int j = i + 1; j < N; ++j, ++p) {
                    if (j < n) {
                        float dx = g[i].x - g[j].x, dy = g[i].y - g[j].y, dz = g[i].z - g[j].z, dw = g[i].w - g[j].w;
                        acc[p] += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                    }
                }
            }
        }
        // block reduction, fixed order: lanes (xor tree) -> warps (ascending) -> CTAs (ascending) -> ranks (ascending)
        int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            float s = warp_sum(acc[p]);
            if (lane == 0)
                sh.warp_partials[warp][p] = s;
        }
        __syncthreads();
        if (threadIdx.x < NP) {
            float s = 0.f;
            for (int w = 0; w < nwarps; ++w)
                s += sh.warp_partials[w][threadIdx.x];
            a.cta_partials[blockIdx.x * kMaxPairs + threadIdx.x] = s;
        }
        __threadfence();
        grid.sync();
        if (blockIdx.x == 0) {
            if (threadIdx.x < NP) {
                float s = 0.f;
                for (unsigned b = 0; b < gridDim.x; ++b)
                    s += ld_volatile_f(a.cta_partials + b * kMaxPairs + threadIdx.x);
                for (int q = 0; q < a.R; ++q)
                    a.mailbox[q][a.rank * kMaxPairs + threadIdx.x] = s;
            }
            fence_sys();
            __syncthreads();
            if (multi)
                signal_all(a, 1);
        }
        if (multi) {
            wait_all(a, 1);
            __syncthreads();
        } else {
            grid.sync();
        }
        // -------- full distance matrix (rank order => identical everywhere) + selection -------- //
        if (threadIdx.x < NP) {
            int i = 0, rest = threadIdx.x;
            while (rest >= N - 1 - i) { // pair index -> (i, j) of the N-padded upper triangle
                rest -= N - 1 - i;
                ++i;
            }
            int const j = i + 1 + rest;
            if (j < n) {
                float s = 0.f;
                for (int q = 0; q < a.R; ++q)
                    s += ld_volatile_f(a.mailbox[a.rank] + q * kMaxPairs + threadIdx.x);
                if (!is_finite(s))
                    s = __int_as_float(0x7f800000);
                sh.dist[i][j] = s;
                sh.dist[j][i] = s;
                if (a.dist_out && blockIdx.x == 0) {
                    a.dist_out[i * n + j] = s;
                    a.dist_out[j * n + i] = s;
                }
            }
        }
        if (threadIdx.x < N)
            sh.dist[threadIdx.x][threadIdx.x] = 0.f;
        __syncthreads();
        if (threadIdx.x < 32) {
            if (a.rule == kKrum)
                select_krum(a, sh);
            else
                select_bulyan(a, sh);
        }
        __syncthreads();
        if (a.info && blockIdx.x == 0 && threadIdx.x < sh.theta) {
            a.info[0] = sh.theta;
            a.info[1 + threadIdx.x] = static_cast<int>(sh.selmask[threadIdx.x]);
        }
        // -------- phase D: aggregate the slice from the staged copy, update, broadcast -------- //
        unsigned needed = 0;
        int const theta = sh.theta;
        for (int k = 0; k < theta; ++k)
            needed |= sh.selmask[k];
        for (long long v = tid; v < len4; v += nthreads) {
            long long const x = a.lo + (v << 2);
            float4 g[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (i < n && ((needed >> i) & 1u))
                    g[i] = a.staging ? *reinterpret_cast<float4 const*>(a.staging + i * len + (v << 2)) : ld_stream_f4(a.grad[i] + x);
                else
                    g[i] = f4_zero();
            }
            float4 out;
            if (a.rule == kKrum) {
                unsigned const mask = sh.selmask[0];
                float4 sum = f4_zero();
#pragma unroll
                for (int i = 0; i < N; ++i)
                    if ((mask >> i) & 1u)
                        sum = f4_add(sum, g[i]);
                float const count = static_cast<float>(a.m);
                out = make_float4(sum.x / count, sum.y / count, sum.z / count, sum.w / count);
            } else {
                float4 inter[N];
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    inter[k] = f4_zero();
                    if (k < theta) {
                        unsigned const mask = sh.selmask[k];
                        float4 sum = f4_zero();
#pragma unroll
                        for (int i = 0; i < N; ++i)
                            if ((mask >> i) & 1u)
                                sum = f4_add(sum, g[i]);
                        float const count = static_cast<float>(sh.selcount[k]);
                        inter[k] = make_float4(sum.x / count, sum.y / count, sum.z / count, sum.w / count);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float vals[N];
#pragma unroll
                    for (int k = 0; k < N; ++k)
                        vals[k] = f4_at(inter[k], c);
                    f4_at(out, c) = coord_averaged_median<N>(vals, theta, a.beta);
                }
            }
            apply_update(a, x, out);
        }
    } else {
        // -------- coordinate-wise rules: one streaming pass -------- //
        for (long long v = tid; v < len4; v += nthreads) {
            long long const x = a.lo + (v << 2);
            float4 g[N];
            bool const in_switch = a.rule == kAverage && a.grad_mc != nullptr;
#pragma unroll
            for (int i = 0; i < N; ++i)
                g[i] = (i < n && !in_switch) ? ld_stream_f4(a.grad[i] + x) : f4_zero();
            float4 out;
            if (a.rule == kAverage) {
                float4 sum = f4_zero();
                if (a.grad_mc) {   // NVLS: the switch adds the same row of every rank; rows of one rank are added here
                    for (int j = 0; j < a.workers_per_rank; ++j)
                        sum = f4_add(sum, multimem_ld_reduce_add_f4(a.grad_mc + j * a.row_stride + x));
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i)
                        if (i < n)
                            sum = f4_add(sum, g[i]);
                }
                float const count = static_cast<float>(n);
                out = make_float4(sum.x / count, sum.y / count, sum.z / count, sum.w / count);
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float vals[N];
#pragma unroll
                    for (int i = 0; i < N; ++i)
                        vals[i] = f4_at(g[i], c);
                    float r;
                    if (a.rule == kAverageNan)
                        r = coord_average_nan<N>(vals, n);
                    else if (a.rule == kMedian)
                        r = coord_median<N>(vals, n);
                    else
                        r = coord_averaged_median<N>(vals, n, a.beta);
                    f4_at(out, c) = r;
                }
            }
            apply_update(a, x, out);
        }
    }

    if (multi) { // exit barrier: all slices have landed on all ranks, gradient buffers are free again
        fence_sys();
        grid.sync();
        if (blockIdx.x == 0) {
            signal_all(a, 2);
            wait_all(a, 2);
        }
    }
}

// ---- small stand-alone kernels (baseline path, attacks, diagnostics) ------------- //
__global__ void sgd_kernel(float* __restrict__ p, float const* __restrict__ g, float lr, long long d) {
    long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
    for (; i + 3 < d; i += stride) {
        float4 pv = *reinterpret_cast<float4 const*>(p + i), gv = *reinterpret_cast<float4 const*>(g + i);
        pv.x -= lr * gv.x; pv.y -= lr * gv.y; pv.z -= lr * gv.z; pv.w -= lr * gv.w;
        *reinterpret_cast<float4*>(p + i) = pv;
    }
}

// Lossy-transport emulation (reference: tf_patches mpi_rendezvous_mgr.patch:814-843): every `chunk`-byte
// datagram of the serialized gradient is lost with probability `rate`; lost chunks become NaN (mode 0),
// zeros (mode 1) or the previous gradient's bytes (mode 2, "CLEVER").
__global__ void drop_chunks_kernel(float* g, float const* previous, long long d, long long chunk_elems, float rate, int mode, unsigned long long seed) {
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (; i < d; i += stride) {
        unsigned long long h = (static_cast<unsigned long long>(i / chunk_elems) + 1) * 0x9E3779B97F4A7C15ull ^ seed;
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        float u = static_cast<float>(h >> 40) * (1.0f / 16777216.0f);
        if (u < rate)
            g[i] = mode == 0 ? __int_as_float(0x7fc00000) : mode == 1 ? 0.f : previous[i];
    }
}

// Order-independent 64-bit checksum of a float buffer (debug: cross-rank parameter equality).
__global__ void checksum_kernel(float const* p, long long d, unsigned long long* out) {
    unsigned long long local = 0;
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (; i < d; i += stride) {
        unsigned long long h = (static_cast<unsigned long long>(__float_as_uint(p[i])) << 20) ^ static_cast<unsigned long long>(i);
        h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
        local += h;
    }
    for (int offset = 16; offset > 0; offset >>= 1)
        local += __shfl_xor_sync(0xffffffffu, local, offset);
    if ((threadIdx.x & 31) == 0)
        atomicAdd(out, local);
}

__global__ void cast_bf16_kernel(float const* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long d) {
    long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
    for (; i + 3 < d; i += stride) {
        float4 v = *reinterpret_cast<float4 const*>(src + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + i) = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
    }
}

template<int N> int launch(GarArgs& a, int max_ctas, cudaStream_t stream) {
    int const threads = N <= 8 ? 512 : 256;
    int device = 0, sms = 0, per_sm = 0;
    AGB_CUDA_OK(cudaGetDevice(&device));
    AGB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    AGB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gar_fused_kernel<N>, threads, 0));
    if (per_sm < 1)
        return 2;
    long long const len4 = (a.hi - a.lo) / 4;
    long long want = (len4 + threads - 1) / threads;
    int grid = sms * per_sm;
    if (max_ctas > 0 && grid > max_ctas)
        grid = max_ctas;
    if (want < grid)
        grid = want < 1 ? 1 : static_cast<int>(want);
    void* params[] = {&a};
    AGB_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gar_fused_kernel<N>), dim3(grid), dim3(threads), params, 0, stream));
    return 0;
}

} // namespace

extern "C" {

char const* agb_op_list() {
    return "gar_fused,gar_max_ctas,sgd,drop_chunks,checksum,cast_bf16";
}

// Wall-clock bound (seconds, 0 = none) of the cross-GPU flag waits of this library's kernels.
int agb_gar_set_flag_timeout(double seconds) {
    return set_flag_timeout(seconds);
}

// Upper bound of the grid the fused kernel may use (to size `cta_partials`: [ctas][120] floats).
int agb_gar_max_ctas() {
    int device = 0, sms = 0;
    if (cudaGetDevice(&device) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
        return 0;
    return sms * 4;
}

// ptrs layout (all device addresses, 0 = absent):
//   [0..16) gradient rows | [16] agg_out | [17] param | [18] slot0 | [19] slot1 | [20] param_mc
//   [21] cta_partials | [22] staging | [23] dist_out | [24] info | [25] grad_mc (multicast address of the gradient matrix)
//   [32..48) param_dst | [48..64) signal | [64..80) mailbox | [80..96) param_bf16_dst
// ints: n f m beta rule R rank opt epoch max_ctas workers_per_rank ; longs: lo hi row_stride ; floats: lr h0 h1 h2
int agb_gar_fused(unsigned long long const* ptrs, int const* ints, long long const* longs, float const* floats, void* stream) {
    GarArgs a{};
    a.n = ints[0]; a.f = ints[1]; a.m = ints[2]; a.beta = ints[3]; a.rule = ints[4];
    a.R = ints[5]; a.rank = ints[6]; a.opt = ints[7]; a.epoch = static_cast<uint32_t>(ints[8]);
    int const max_ctas = ints[9];
    a.lo = longs[0]; a.hi = longs[1];
    a.lr = floats[0]; a.h0 = floats[1]; a.h1 = floats[2]; a.h2 = floats[3];
    if (a.n < 1 || a.n > kMaxWorkers || a.R < 1 || a.R > kMaxRanks || a.rank < 0 || a.rank >= a.R)
        return 100;
    if ((a.lo & 3) || (a.hi & 3) || a.hi < a.lo)
        return 101;
    if (a.rule < 0 || a.rule > kBulyan)
        return 102;
    if ((a.rule == kKrum || a.rule == kBulyan) && (a.n - a.f - 2 < 1 || a.m < 1 || a.m > a.n))
        return 103;
    if (a.rule == kBulyan && (a.n < 4 * a.f + 3 || a.beta != a.n - 4 * a.f - 2 || a.m < a.n - 2 * a.f - 2))
        return 104;
    if (a.rule == kAveragedMedian && (a.beta < 1 || a.beta > a.n))
        return 105;
    for (int i = 0; i < a.n; ++i)
        a.grad[i] = reinterpret_cast<float const*>(ptrs[i]);
    a.agg_out = reinterpret_cast<float*>(ptrs[16]);
    a.param = reinterpret_cast<float*>(ptrs[17]);
    a.slot0 = reinterpret_cast<float*>(ptrs[18]);
    a.slot1 = reinterpret_cast<float*>(ptrs[19]);
    a.param_mc = reinterpret_cast<float*>(ptrs[20]);
    a.cta_partials = reinterpret_cast<float*>(ptrs[21]);
    a.staging = reinterpret_cast<float*>(ptrs[22]);
    a.dist_out = reinterpret_cast<float*>(ptrs[23]);
    a.info = reinterpret_cast<int*>(ptrs[24]);
    a.grad_mc = reinterpret_cast<float const*>(ptrs[25]);
    a.workers_per_rank = ints[10];
    a.row_stride = longs[2];
    for (int q = 0; q < a.R; ++q) {
        a.param_dst[q] = reinterpret_cast<float*>(ptrs[32 + q]);
        a.signal[q] = reinterpret_cast<uint32_t*>(ptrs[48 + q]);
        a.mailbox[q] = reinterpret_cast<float*>(ptrs[64 + q]);
        a.param_bf16_dst[q] = reinterpret_cast<__nv_bfloat16*>(ptrs[80 + q]);
    }
    if (a.opt != kNone && (!a.param || (!a.param_mc && !a.param_dst[0])))
        return 106;
    if ((a.opt == kAdam || a.opt == kRmsprop || a.opt == kAdadelta) && (!a.slot0 || !a.slot1))
        return 107;
    if (a.opt == kAdagrad && !a.slot0)
        return 107;
    if ((a.rule == kKrum || a.rule == kBulyan) && (!a.cta_partials || !a.mailbox[0]))
        return 108;
    if (a.R > 1 && !a.signal[0])
        return 109;
    if (a.hi == a.lo && a.R == 1)
        return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return a.n <= 8 ? launch<8>(a, max_ctas, s) : launch<16>(a, max_ctas, s);
}

int agb_sgd(void* p, void const* g, float lr, long long d, void* stream) {
    if (d & 3)
        return 101;
    int blocks = static_cast<int>((d / 4 + 255) / 256);
    if (blocks > 148 * 8)
        blocks = 148 * 8;
    if (blocks > 0)
        sgd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float*>(p), static_cast<float const*>(g), lr, d);
    AGB_CUDA_OK(cudaGetLastError());
    return 0;
}

int agb_drop_chunks(void* g, void const* previous, long long d, long long chunk_bytes, float rate, int mode, unsigned long long seed, void* stream) {
    if (chunk_bytes < 4 || (mode == 2 && !previous))
        return 101;
    int blocks = static_cast<int>((d + 255) / 256);
    if (blocks > 148 * 8)
        blocks = 148 * 8;
    if (blocks > 0)
        drop_chunks_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float*>(g), static_cast<float const*>(previous), d, chunk_bytes / 4, rate, mode, seed);
    AGB_CUDA_OK(cudaGetLastError());
    return 0;
}

int agb_checksum(void const* p, long long d, void* out, void* stream) {
    AGB_CUDA_OK(cudaMemsetAsync(out, 0, 8, static_cast<cudaStream_t>(stream)));
    int blocks = static_cast<int>((d + 255) / 256);
    if (blocks > 148 * 4)
        blocks = 148 * 4;
    if (blocks > 0)
        checksum_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float const*>(p), d, static_cast<unsigned long long*>(out));
    AGB_CUDA_OK(cudaGetLastError());
    return 0;
}

int agb_cast_bf16(void const* src, void* dst, long long d, void* stream) {
    if (d & 3)
        return 101;
    int blocks = static_cast<int>((d / 4 + 255) / 256);
    if (blocks > 148 * 8)
        blocks = 148 * 8;
    if (blocks > 0)
        cast_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float const*>(src), static_cast<__nv_bfloat16*>(dst), d);
    AGB_CUDA_OK(cudaGetLastError());
    return 0;
}

} // extern "C"
