// SHA-256 tree digest of a device buffer: the cryptographic digest that gradient authentication signs (`parallel/signing.py`).
//
// The reference's hardened transport signs every worker -> parameter-server message with ed25519
// (`tf_patches/patches/mpi_rendezvous_mgr.patch:514-523,777-781`); here the signed unit is the digest of a published gradient slice,
// recomputed by the consumer through the peer mapping. A 100 MB slice cannot be hashed serially on a GPU, so the digest is a tree:
//
//   level 0: the buffer is cut into 1024-byte leaves; node i = SHA-256( level(u32) | 0(u32) | i(u64) | leaf bytes )
//   level k: the 32-byte nodes of level k-1, concatenated, are cut into 1024-byte groups and hashed the same way
//   ... until one node is left: the digest. The (level, index) header pins every node to its place in the tree.
//
// One thread hashes one leaf (<= 17 compression rounds); level 0 of a 100 MB slice is 100 000 independent threads.

#include <cstdint>
#include <cuda_runtime.h>

#include <agb_device.cuh>

namespace {

constexpr int kLeafBytes = 1024;
constexpr int kThreads = 128;

__constant__ uint32_t kRound[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74,
    0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d,
    0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e,
    0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5,
    0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) {
    return __funnelshift_r(x, x, n);
}

// One SHA-256 compression of the 16 big-endian words in `w` (destroyed: used as the rolling message schedule).
__device__ void compress(uint32_t (&h)[8], uint32_t (&w)[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], k = h[7];
#pragma unroll
    for (int t = 0; t < 64; ++t) {
        if (t >= 16) {
            uint32_t const w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
            uint32_t const s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3), s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            w[t & 15] += s0 + w[(t + 9) & 15] + s1;
        }
        uint32_t const t1 = k + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + kRound[t] + w[t & 15];
        uint32_t const t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        k = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += k;
}

// `in`: nbytes (multiple of 4) bytes; `out`: ceil(nbytes / 1024) nodes of 32 bytes (at least one: the empty buffer has one leaf).
__global__ void __launch_bounds__(kThreads) sha256_level_kernel(uint32_t const* __restrict__ in, long long nbytes, uint32_t* __restrict__ out, long long nodes, uint32_t level) {
    long long const node = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
    if (node >= nodes)
        return;
    long long const begin = node * kLeafBytes;
    int const len = static_cast<int>(nbytes - begin < kLeafBytes ? (nbytes - begin > 0 ? nbytes - begin : 0) : kLeafBytes);   // bytes of this leaf
    int const words = len >> 2;
    uint32_t const* src = in + (begin >> 2);
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t w[16];
    // the byte stream: 16 header bytes, `len` payload bytes (memory order), 0x80, zeros, 64-bit big-endian bit count; SHA-256 packs
    // the stream into big-endian words, so little-endian memory words are byte-swapped on the way in
    unsigned long long const bits = (16ull + static_cast<unsigned long long>(len)) * 8ull;
    int const marker = 4 + words;                               // word position of the 0x80 marker
    int const blocks = (marker + 1 + 2 + 15) / 16;
    bool const aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    for (int blk = 0; blk < blocks; ++blk) {
        int const first = blk * 16;
        if (blk > 0 && first + 16 <= marker && aligned) {       // 16 payload words: four 16-byte loads ((first - 4) * 4 is a multiple of 16)
            uint4 const* vec = reinterpret_cast<uint4 const*>(src + (first - 4));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 const v = vec[q];
                w[4 * q + 0] = __byte_perm(v.x, 0, 0x0123);
                w[4 * q + 1] = __byte_perm(v.y, 0, 0x0123);
                w[4 * q + 2] = __byte_perm(v.z, 0, 0x0123);
                w[4 * q + 3] = __byte_perm(v.w, 0, 0x0123);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                int const pos = first + j;
                uint32_t v = 0;
                if (pos >= 4 && pos < marker)
                    v = __byte_perm(src[pos - 4], 0, 0x0123);
                else if (pos == marker)
                    v = 0x80000000u;
                w[j] = v;
            }
            if (blk == 0) {
                w[0] = __byte_perm(level, 0, 0x0123);           // header: level (u32 LE), 0 (u32), node index (u64 LE)
                w[1] = 0;
                w[2] = __byte_perm(static_cast<uint32_t>(node), 0, 0x0123);
                w[3] = __byte_perm(static_cast<uint32_t>(static_cast<unsigned long long>(node) >> 32), 0, 0x0123);
            }
            if (blk == blocks - 1) {
                w[14] = static_cast<uint32_t>(bits >> 32);
                w[15] = static_cast<uint32_t>(bits);
            }
        }
        compress(h, w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        out[node * 8 + j] = __byte_perm(h[j], 0, 0x0123);                  // digest bytes in SHA-256's own (big-endian) order
}

} // namespace

extern "C" {

// Scratch requirement of `agb_sha256_tree` for a buffer of `nbytes` bytes.
long long agb_sha256_scratch_bytes(long long nbytes) {
    long long const level0 = nbytes <= 0 ? 1 : (nbytes + kLeafBytes - 1) / kLeafBytes;
    long long const level1 = (level0 * 32 + kLeafBytes - 1) / kLeafBytes;
    return (level0 + level1) * 32 + 64;
}

// digest[32] = SHA-256 tree digest of data[0, nbytes); nbytes % 4 == 0, data 4-byte aligned (16 for speed). `scratch`: see above.
int agb_sha256_tree(void const* data, long long nbytes, void* scratch, void* digest, void* stream) {
    if (nbytes < 0 || (nbytes & 3) || (reinterpret_cast<uintptr_t>(data) & 3))
        return 121;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    long long const level0 = nbytes == 0 ? 1 : (nbytes + kLeafBytes - 1) / kLeafBytes;
    uint32_t* region[2] = {static_cast<uint32_t*>(scratch), static_cast<uint32_t*>(scratch) + level0 * 8};   // level 0, 2, ... / level 1, 3, ... outputs
    uint32_t const* in = static_cast<uint32_t const*>(data);
    long long in_bytes = nbytes, nodes = level0;
    for (uint32_t level = 0;; ++level) {
        uint32_t* out = nodes == 1 ? static_cast<uint32_t*>(digest) : region[level & 1];
        sha256_level_kernel<<<static_cast<unsigned>((nodes + kThreads - 1) / kThreads), kThreads, 0, s>>>(in, in_bytes, out, nodes, level);
        AGB_CUDA_OK(cudaGetLastError());
        if (nodes == 1)
            break;
        in = out;
        in_bytes = nodes * 32;
        nodes = (in_bytes + kLeafBytes - 1) / kLeafBytes;
    }
    return 0;
}

} // extern "C"
