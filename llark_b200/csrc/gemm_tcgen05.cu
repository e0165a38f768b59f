// llark_b200 GEMM core: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (+ R[M,N]),  fp32 accumulate in TMEM.
//
// Replaces every cuBLAS `addmm`/`Linear` call on the reference hot path:
// reference call sites m2t/models/llamav2.py:133 (mm_projector), :224 (LlamaModel q/k/v/o/gate/up/down),
// :312 (lm_head), their autograd dgrad/wgrad, and jukebox/main.py:108 (prior c_attn/c_proj/mlp).
//
// Design (one CTA per SM, 256 threads):
//   warp 0   : TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle, mbarrier complete_tx)
//   warp 1   : MMA issuer    (one lane issues tcgen05.mma kind::f16 / kind::tf32, M=128, N=256, K=32 bytes)
//   warp 2   : TMEM allocator (512 columns = two 128x256 fp32 accumulator stages)
//   warps 4-7: epilogue       (tcgen05.ld 32x32b -> registers -> bias/residual -> bf16/fp32 -> 16 B global stores)
// The smem ring has kStages stages of (A 16 KB + B 32 KB); the two TMEM accumulator stages let the MMA
// of tile i+1 overlap the epilogue of tile i. Operands may be K-major or MN-major (dgrad/wgrad read the
// weights / activations in place, no transposes), selected per operand through the UMMA descriptors.
#include "common.cuh"
#include "../../include/llark_b200.h"
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace llark {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_KB = 128;  // bytes of K per smem row == swizzle width
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kThreads = 256;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_KB;  // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_KB;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
// Epilogue staging: every epilogue warp owns a [32 rows][32 + 4 pad fp32] patch of shared memory through which a 32-column
// accumulator chunk is transposed from "one row per thread" (what tcgen05.ld delivers) to "a quarter-warp per row", so
// that the residual loads and the C stores of a warp instruction cover whole 128 B (fp32) / 64 B (bf16) row segments
// instead of 32 different rows (see epilogue_tile_staged).
constexpr int EPI_LD = 36;                        // floats per staged row; 36 = 32 + 4 keeps float4 rows conflict-free
constexpr int EPI_WARP_BYTES = 32 * EPI_LD * 4;   // 4608 B per epilogue warp
constexpr int EPI_STAGE_BYTES = 4 * EPI_WARP_BYTES;
constexpr int SMEM_BYTES = kStages * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;

struct GemmParams {
    int M, N, K;
    void* C;
    long long ldc;
    int c_fp32;
    const void* bias;  // bf16 [N] (or fp32 when bias_fp32)
    int bias_fp32;
    const void* R;  // residual, same dtype as C
    long long ldr;
    int num_m_blocks, num_n_blocks, num_k_blocks;
    int elem_bytes;  // 2 (bf16) or 4 (tf32)
    int a_tma4, b_tma4;  // MN-major operand loaded by ONE 4-D TMA box per stage (atom order: 64 MN, 8 K, chunks, K-groups)
    int* tile_counter;   // CTA-pair kernel: dynamic tile scheduler counter (zero at launch, reset by the last fetch)
    int debug;       // LLARK_GEMM_DEBUG bits (perf experiments only): 1 = split K-major boxes in 64-row halves,
                     // 2 = producer issues no TMA (MMA runs on stale smem: isolates the UMMA rate)
    // fused epilogue (LlarkGemmEpilogue): 0 none, 1 quick-GELU, 2 RoPE on the first epi_i1 columns (head_dim 128,
    // position = row % epi_i0, epi_a / epi_b = cos / sin bf16 [S,128]), 3 SwiGLU backward (epi_a = gu bf16 [M, epi_ld]:
    // gate | up halves N apart; C receives dgate at column c and dup at column N + c)
    int staged;      // 1 = staged (coalesced) epilogue, 0 = direct row-per-thread stores (see epilogue_dispatch)
    int group_m;     // m-blocks per rasterisation group (tile_coord): sized on the host so a group's A panel stays in L2
    int epi;
    const void* epi_a;
    const void* epi_b;
    long long epi_ld;
    int epi_i0, epi_i1;
};

__device__ __forceinline__ void ld8_bf16(const __nv_bfloat16* ptr, float (&f)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(ptr);
    const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* ptr, const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(ptr) = o;
}

// SM100 shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct TileCoord {
    int m_blk, n_blk;
};
// Grouped rasterisation: 16 m-blocks share a sweep over n so that the ~148 tiles in flight reuse
// A/B panels out of L2.
__device__ __forceinline__ TileCoord tile_coord(int tile, int num_m_blocks, int num_n_blocks, int GROUP_M = 16) {
    const int tiles_per_group = GROUP_M * num_n_blocks;
    const int group = tile / tiles_per_group;
    const int first_m = group * GROUP_M;
    const int group_m = min(GROUP_M, num_m_blocks - first_m);
    const int in_group = tile - group * tiles_per_group;
    TileCoord t;
    t.m_blk = first_m + in_group % group_m;
    t.n_blk = in_group / group_m;
    return t;
}

// One 32-column chunk of one accumulator row: bias / residual, conversion, 16-byte stores (N tail masked).
template <int EPI>
__device__ __forceinline__ void epilogue_store_chunk(const GemmParams& p, long long row, int col0,
                                     const uint32_t (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = col0 + j * 8;
        if (col >= p.N) break;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]);
        const bool full8 = (col + 8 <= p.N);
        if (p.bias) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (full8 || col + e < p.N)
                    f[e] += p.bias_fp32 ? reinterpret_cast<const float*>(p.bias)[col + e]
                                        : __bfloat162float(
                                              reinterpret_cast<const __nv_bfloat16*>(p.bias)[col + e]);
        }
        if constexpr (EPI == 1) {
            // quick-GELU of the prior MLP (x * sigmoid(1.702 x)), same formula as llark_quick_gelu
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = f[e] / (1.f + __expf(-1.702f * f[e]));
        } else if constexpr (EPI == 3) {
            // SwiGLU backward fused into the down-projection dgrad: the accumulator is d(act) (rounded to bf16 as
            // the unfused GEMM would store it); same arithmetic as swiglu_bwd_kernel (llama_ops.cu)
            const __nv_bfloat16* gptr = reinterpret_cast<const __nv_bfloat16*>(p.epi_a) + row * p.epi_ld + col;
            __nv_bfloat16* cptr = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col;
            float g[8], u[8], dg[8], du[8];
            ld8_bf16(gptr, g);
            ld8_bf16(gptr + p.N, u);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = bf16_round(f[e]);
                const float sg = __fdividef(1.f, 1.f + __expf(-g[e]));
                const float silu = g[e] * sg;
                du[e] = d * bf16_round(silu);
                dg[e] = d * u[e] * (sg * (1.f + g[e] * (1.f - sg)));
            }
            st8_bf16(cptr, dg);
            st8_bf16(cptr + p.N, du);
            continue;
        }
        if (p.c_fp32) {
            float* cptr = reinterpret_cast<float*>(p.C) + row * p.ldc + col;
            if (p.R) {
                const float* rptr = reinterpret_cast<const float*>(p.R) + row * p.ldr + col;
                if (full8) {
                    const float4 r0 = *reinterpret_cast<const float4*>(rptr);
                    const float4 r1 = *reinterpret_cast<const float4*>(rptr + 4);
                    f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
                    f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
                } else {
                    for (int e = 0; e < 8; ++e)
                        if (col + e < p.N) f[e] += rptr[e];
                }
            }
            if (full8) {
                *reinterpret_cast<float4*>(cptr) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(cptr + 4) = make_float4(f[4], f[5], f[6], f[7]);
            } else {
                for (int e = 0; e < 8; ++e)
                    if (col + e < p.N) cptr[e] = f[e];
            }
        } else {
            __nv_bfloat16* cptr = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col;
            if (p.R) {
                const __nv_bfloat16* rptr = reinterpret_cast<const __nv_bfloat16*>(p.R) + row * p.ldr + col;
                if (full8) {
                    const uint4 r = *reinterpret_cast<const uint4*>(rptr);
                    const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c2 = unpack_bf16x2(r.z),
                                 d = unpack_bf16x2(r.w);
                    f[0] += a.x; f[1] += a.y; f[2] += b.x; f[3] += b.y;
                    f[4] += c2.x; f[5] += c2.y; f[6] += d.x; f[7] += d.y;
                } else {
                    for (int e = 0; e < 8; ++e)
                        if (col + e < p.N) f[e] += __bfloat162float(rptr[e]);
                }
            }
            if (full8) {
                uint4 o;
                o.x = pack_bf16x2(f[0], f[1]);
                o.y = pack_bf16x2(f[2], f[3]);
                o.z = pack_bf16x2(f[4], f[5]);
                o.w = pack_bf16x2(f[6], f[7]);
                *reinterpret_cast<uint4*>(cptr) = o;
            } else {
                for (int e = 0; e < 8; ++e)
                    if (col + e < p.N) cptr[e] = __float2bfloat16_rn(f[e]);
            }
        }
    }
}

// All 256 accumulator columns of one row (this thread's TMEM lane): 8 chunks of 32 columns through
// epilogue_store_chunk, or -- for the q and k columns of a fused-RoPE QKV projection -- pairs of chunks 64 columns
// apart rotated as HF apply_rotary_pos_emb does in bf16: x = bf16(acc) (the Linear output),
// out = bf16(bf16(x*cos) + bf16(rotate_half(x)*sin)), identical to llark_rope applied after an unfused GEMM.
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, long long row, bool row_ok, int ncol0,
                                              uint32_t taddr) {
    if (EPI == 2 && ncol0 < p.epi_i1) {
        const int spos = (int)(row % p.epi_i0);
        const __nv_bfloat16* cs = reinterpret_cast<const __nv_bfloat16*>(p.epi_a) + (long long)spos * 128;
        const __nv_bfloat16* sn = reinterpret_cast<const __nv_bfloat16*>(p.epi_b) + (long long)spos * 128;
#pragma unroll 1
        for (int hc = 0; hc < 4; ++hc) {
            const int clo = (hc >> 1) * 128 + (hc & 1) * 32;  // tile-local column of the low-half chunk
            if (ncol0 + clo >= p.N) break;                    // warp-uniform (N % 128 == 0 is checked on the host)
            uint32_t lo[32], hi[32];
            tmem_ld_32x32b_x32(taddr + clo, lo);
            tmem_ld_32x32b_x32(taddr + clo + 64, hi);
            tmem_ld_wait();
            if (!row_ok) continue;
            __nv_bfloat16* cptr = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + ncol0 + clo;
            const int i0 = (hc & 1) * 32;  // index inside the half head
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float c1[8], s1[8], c2[8], s2[8], o1[8], o2[8];
                ld8_bf16(cs + i0 + j * 8, c1);
                ld8_bf16(sn + i0 + j * 8, s1);
                ld8_bf16(cs + 64 + i0 + j * 8, c2);
                ld8_bf16(sn + 64 + i0 + j * 8, s2);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float x1 = bf16_round(__uint_as_float(lo[j * 8 + e]));
                    const float x2 = bf16_round(__uint_as_float(hi[j * 8 + e]));
                    o1[e] = bf16_round(x1 * c1[e]) + bf16_round(-x2 * s1[e]);
                    o2[e] = bf16_round(x2 * c2[e]) + bf16_round(x1 * s2[e]);
                }
                st8_bf16(cptr + j * 8, o1);
                st8_bf16(cptr + 64 + j * 8, o2);
            }
        }
        return;
    }
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int col0 = ncol0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c * 32, v);
        tmem_ld_wait();
        if (!row_ok) continue;
        epilogue_store_chunk<EPI>(p, row, col0, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Staged (coalesced) epilogue. tcgen05.ld hands every thread 32 consecutive columns of ONE accumulator row, so direct
// 16 B stores make each warp instruction touch 32 different rows (32 L2 sectors, half-filled) -- measured at ~35 us per
// 256x256 fp32 tile with a residual, more than the MMA time of the tile for K <= 4800. Here the chunk goes through the
// warp's shared-memory patch and comes back with a quarter-warp (fp32: 8 lanes x 16 B = 128 B) or an eighth-warp (bf16:
// 4 lanes x 16 B = 64 B) per row: full sectors, 8x / 4x fewer memory transactions per instruction. The residual (or the
// gate|up operand of the SwiGLU backward) of chunk c+1 is requested before chunk c is finished, so its latency overlaps
// the TMEM read and the transpose. Every output element is still produced by exactly the arithmetic of
// epilogue_store_chunk (same order of bias / activation / residual / rounding), and the thread that reads R[row, col]
// is the one that writes C[row, col], so C may alias R (in-place accumulation).
template <int EPI>
__device__ __forceinline__ void epi_staged_f32(const GemmParams& p, long long row0, int ncol0, uint32_t taddr, float* stg,
                                               int lane) {
    const int rsub = lane >> 3, cq = (lane & 7) * 4;
    float* const C = reinterpret_cast<float*>(p.C);
    const float* const R = reinterpret_cast<const float*>(p.R);
    const int nchunks = min(BLOCK_N / 32, (p.N - ncol0 + 31) / 32);  // warp-uniform
    float4 rc[8], rn[8];
    auto load_res = [&](int col0, float4 (&r)[8]) {
        const int col = col0 + cq;
        const int nc = p.N - col;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long row = row0 + i * 4 + rsub;
            r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= p.M || nc <= 0) continue;
            const float* rp = R + row * p.ldr + col;
            if (nc >= 4) {
                r[i] = *reinterpret_cast<const float4*>(rp);
            } else {
                r[i].x = rp[0];
                if (nc > 1) r[i].y = rp[1];
                if (nc > 2) r[i].z = rp[2];
            }
        }
    };
    if (R) load_res(ncol0, rc);
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        const int col0 = ncol0 + c * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c * 32, v);
        tmem_ld_wait();
        float* srow = stg + lane * EPI_LD;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(srow + 4 * j) =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        if (R && c + 1 < nchunks) load_res(col0 + 32, rn);
        const int col = col0 + cq;
        const int nc = p.N - col;
        float b[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias && nc > 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (e < nc)
                    b[e] = p.bias_fp32 ? reinterpret_cast<const float*>(p.bias)[col + e]
                                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[col + e]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + rsub;
            const long long row = row0 + r;
            const float4 a = *reinterpret_cast<const float4*>(stg + r * EPI_LD + cq);
            if (row >= p.M || nc <= 0) continue;
            float f[4] = {a.x + b[0], a.y + b[1], a.z + b[2], a.w + b[3]};
            if constexpr (EPI == 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) f[e] = f[e] / (1.f + __expf(-1.702f * f[e]));
            }
            if (R) {
                f[0] += rc[i].x; f[1] += rc[i].y; f[2] += rc[i].z; f[3] += rc[i].w;
            }
            float* cp = C + row * p.ldc + col;
            if (nc >= 4) {
                *reinterpret_cast<float4*>(cp) = make_float4(f[0], f[1], f[2], f[3]);
            } else {
                for (int e = 0; e < nc; ++e) cp[e] = f[e];
            }
        }
        __syncwarp();  // the patch is rewritten by the next chunk
        if (R) {
#pragma unroll
            for (int i = 0; i < 8; ++i) rc[i] = rn[i];
        }
    }
}

template <int EPI>
__device__ __forceinline__ void epi_staged_bf16(const GemmParams& p, long long row0, int ncol0, uint32_t taddr, float* stg,
                                                int lane) {
    const int rsub = lane >> 2, cq = (lane & 3) * 8;
    __nv_bfloat16* const C = reinterpret_cast<__nv_bfloat16*>(p.C);
    // operand streamed alongside the accumulator: the residual R, or (EPI 3) gu = gate | up of the SwiGLU backward
    const __nv_bfloat16* const X = EPI == 3 ? reinterpret_cast<const __nv_bfloat16*>(p.epi_a)
                                            : reinterpret_cast<const __nv_bfloat16*>(p.R);
    const long long ldx = EPI == 3 ? p.epi_ld : p.ldr;
    const int nchunks = min(BLOCK_N / 32, (p.N - ncol0 + 31) / 32);  // warp-uniform
    uint4 xc[4], xn[4], uc[4], un[4];  // uc / un: the `up` half (EPI 3 only)
    auto load_x = [&](int col0, uint4 (&x)[4], uint4 (&u)[4]) {
        const int col = col0 + cq;
        const int nc = p.N - col;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long row = row0 + i * 8 + rsub;
            x[i] = make_uint4(0u, 0u, 0u, 0u);
            if (EPI == 3) u[i] = make_uint4(0u, 0u, 0u, 0u);
            if (row >= p.M || nc <= 0) continue;
            const __nv_bfloat16* xp = X + row * ldx + col;
            if (nc >= 8) {
                x[i] = *reinterpret_cast<const uint4*>(xp);
                if (EPI == 3) u[i] = *reinterpret_cast<const uint4*>(xp + p.N);
            } else {  // N tail (never taken by EPI 3: N % 8 == 0 is checked on the host)
                __nv_bfloat16 t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = e < nc ? xp[e] : __float2bfloat16_rn(0.f);
                x[i] = *reinterpret_cast<const uint4*>(t);
            }
        }
    };
    if (X) load_x(ncol0, xc, uc);
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        const int col0 = ncol0 + c * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c * 32, v);
        tmem_ld_wait();
        float* srow = stg + lane * EPI_LD;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(srow + 4 * j) =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        if (X && c + 1 < nchunks) load_x(col0 + 32, xn, un);
        const int col = col0 + cq;
        const int nc = p.N - col;
        float b[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) b[e] = 0.f;
        if (p.bias && nc > 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (e < nc)
                    b[e] = p.bias_fp32 ? reinterpret_cast<const float*>(p.bias)[col + e]
                                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[col + e]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + rsub;
            const long long row = row0 + r;
            const float4 a0 = *reinterpret_cast<const float4*>(stg + r * EPI_LD + cq);
            const float4 a1 = *reinterpret_cast<const float4*>(stg + r * EPI_LD + cq + 4);
            if (row >= p.M || nc <= 0) continue;
            float f[8] = {a0.x + b[0], a0.y + b[1], a0.z + b[2], a0.w + b[3],
                          a1.x + b[4], a1.y + b[5], a1.z + b[6], a1.w + b[7]};
            __nv_bfloat16* cp = C + row * p.ldc + col;
            if constexpr (EPI == 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = f[e] / (1.f + __expf(-1.702f * f[e]));
            } else if constexpr (EPI == 3) {
                // same arithmetic as the EPI 3 branch of epilogue_store_chunk / swiglu_bwd_kernel (llama_ops.cu)
                const float2 g0 = unpack_bf16x2(xc[i].x), g1 = unpack_bf16x2(xc[i].y), g2 = unpack_bf16x2(xc[i].z),
                             g3 = unpack_bf16x2(xc[i].w);
                const float2 u0 = unpack_bf16x2(uc[i].x), u1 = unpack_bf16x2(uc[i].y), u2 = unpack_bf16x2(uc[i].z),
                             u3 = unpack_bf16x2(uc[i].w);
                const float g[8] = {g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, g3.x, g3.y};
                const float u[8] = {u0.x, u0.y, u1.x, u1.y, u2.x, u2.y, u3.x, u3.y};
                float dg[8], du[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float d = bf16_round(f[e]);
                    const float sg = __fdividef(1.f, 1.f + __expf(-g[e]));
                    const float silu = g[e] * sg;
                    du[e] = d * bf16_round(silu);
                    dg[e] = d * u[e] * (sg * (1.f + g[e] * (1.f - sg)));
                }
                st8_bf16(cp, dg);
                st8_bf16(cp + p.N, du);
                continue;
            }
            if (EPI != 3 && X) {
                const float2 r0 = unpack_bf16x2(xc[i].x), r1 = unpack_bf16x2(xc[i].y), r2 = unpack_bf16x2(xc[i].z),
                             r3 = unpack_bf16x2(xc[i].w);
                f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
                f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
            }
            if (nc >= 8) {
                st8_bf16(cp, f);
            } else {
                for (int e = 0; e < nc; ++e) cp[e] = __float2bfloat16_rn(f[e]);
            }
        }
        __syncwarp();  // the patch is rewritten by the next chunk
        if (X) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xc[i] = xn[i];
                if (EPI == 3) uc[i] = un[i];
            }
        }
    }
}

// Epilogue of one 128 x 256 accumulator tile for the warp owning TMEM lanes [32 ew, 32 ew + 32): rows row0 .. row0 + 31.
// The RoPE columns of a fused-QKV projection keep the row-per-thread path (rotation partners are 64 columns apart).
// p.staged (chosen on the host, llark_gemm_ex) selects the staged path where the epilogue is the longer leg: a residual
// with an fp32 tile or a short K, and the SwiGLU backward. Plain output at long K stays direct: its epilogue already hides
// behind the next tile's MMAs, and the staging patch's extra 256 KB of shared-memory traffic per tile slows the operand
// pipeline by 3-8 % (profiles/r2c_epilogue_ab.md, profiles/r2k_epilogue_probe.md).
template <int EPI>
__device__ __forceinline__ void epilogue_dispatch(const GemmParams& p, long long row0, int lane, int ncol0, uint32_t taddr,
                                                  float* stg) {
    if ((EPI == 2 && ncol0 < p.epi_i1) || !p.staged) {
        const long long row = row0 + lane;
        epilogue_tile<EPI>(p, row, row < p.M, ncol0, taddr);
        return;
    }
    if (p.c_fp32) {
        if constexpr (EPI == 0 || EPI == 1) epi_staged_f32<EPI>(p, row0, ncol0, taddr, stg, lane);
    } else {
        epi_staged_bf16<EPI>(p, row0, ncol0, taddr, stg, lane);
    }
}

template <int A_MN, int B_MN, int TF32, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + kAccStages;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);
    float* epi_stage = reinterpret_cast<float*>(smem + kStages * STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_blocks * p.num_n_blocks;
    const int elems_per_row = BLOCK_KB / p.elem_bytes;  // K elements per 128 B smem row (64 bf16 / 32 tf32)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == 2) {
        tmem_alloc(tmem_base_slot, kAccStages * BLOCK_N);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const TileCoord tc = tile_coord(tile, p.num_m_blocks, p.num_n_blocks, p.group_m);
            const int m0 = tc.m_blk * BLOCK_M, n0 = tc.n_blk * BLOCK_N;
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                uint8_t* sb = sa + A_STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                const int k0 = kb * elems_per_row;
                if (A_MN == 0) {
                    tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);  // box {K=128B, 128 rows}
                } else {
                    // A stored [K][M]: boxes of {64 M-elements (128 B), elems_per_row K rows}
                    constexpr int MN_PER_BOX_BYTES = 128;
                    const int mn_per_box = MN_PER_BOX_BYTES / p.elem_bytes;
                    const int box_bytes = elems_per_row * 128;
                    for (int j = 0; j < BLOCK_M / mn_per_box; ++j)
                        tma_load_2d(sa + j * box_bytes, &tma_a, &full_bar[stage], m0 + j * mn_per_box, k0);
                }
                if (B_MN == 0) {
                    tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);  // box {K=128B, 256 rows}
                } else {
                    const int mn_per_box = 128 / p.elem_bytes;
                    const int box_bytes = elems_per_row * 128;
                    for (int j = 0; j < BLOCK_N / mn_per_box; ++j)
                        tma_load_2d(sb + j * box_bytes, &tma_b, &full_bar[stage], n0 + j * mn_per_box, k0);
                }
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer =====================
        // Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 @4, a/b format @7/@10
        // (BF16 = 1, TF32 = 2), a_major @15, b_major @16, N>>3 @17, M>>4 @24.
        const uint32_t fmt = TF32 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                               ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
        // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused.
        // MN-major SW128: 128 B of MN per row, 8 K-rows per 1024 B group (SBO), next 128 B MN chunk LBO apart.
        const uint32_t mn_box_bytes = elems_per_row * 128;
        const uint32_t a_lbo = A_MN ? mn_box_bytes : 0, b_lbo = B_MN ? mn_box_bytes : 0;
        // one UMMA consumes 32 bytes of K: K-major -> +32 B inside the swizzle row; MN-major -> +(32/elem) rows
        const uint32_t k_rows_per_mma = 32 / p.elem_bytes;
        const uint32_t a_kstep = A_MN ? (k_rows_per_mma * 128) >> 4 : (32 >> 4);
        const uint32_t b_kstep = B_MN ? (k_rows_per_mma * 128) >> 4 : (32 >> 4);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BLOCK_N;
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t sb = sa + A_STAGE_BYTES;
                const uint64_t adesc = make_smem_desc(sa, a_lbo, 1024);
                const uint64_t bdesc = make_smem_desc(sb, b_lbo, 1024);
#pragma unroll
                for (int k = 0; k < BLOCK_KB / 32; ++k) {
                    const uint64_t ad = adesc + (uint64_t)(k * a_kstep);
                    const uint64_t bd = bdesc + (uint64_t)(k * b_kstep);
                    if (TF32)
                        umma_tf32(tmem_d, ad, bd, idesc, (kb | k) != 0);
                    else
                        umma_f16(tmem_d, ad, bd, idesc, (kb | k) != 0);
                }
                umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(&tmem_full[as]);  // accumulator complete -> epilogue
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;  // TMEM lane quadrant
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const TileCoord tc = tile_coord(tile, p.num_m_blocks, p.num_n_blocks, p.group_m);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const long long row0 = (long long)tc.m_blk * BLOCK_M + ew * 32;
            epilogue_dispatch<EPI>(p, row0, lane, tc.n_blk * BLOCK_N, tmem_base + ((uint32_t)(ew * 32) << 16) + as * BLOCK_N,
                                   epi_stage + ew * (EPI_WARP_BYTES / 4));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kAccStages * BLOCK_N);
    }
}

// =====================================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile. CTA r of the
// pair loads A rows [m0 + 128 r, +128) and B rows (N) [n0 + 128 r, +128) into its own shared memory; the leader
// (rank 0) issues tcgen05.mma.cta_group::2 (M = 256, N = 256) which reads both CTAs' operands, and each CTA's
// TMEM receives its 128 accumulator rows x 256 columns. Per SM the shared-memory traffic drops from
// 96 (TMA writes) + 96 (UMMA reads) B/clk to 64 + 64 B/clk, which is what lets the tensor pipe stay busy.
constexpr int kStages2 = 6;
constexpr int kSched = 4;  // depth of the tile-index ring of the dynamic scheduler
constexpr int B2_STAGE_BYTES = 128 * BLOCK_KB;  // this CTA's half of B: 16 KB
constexpr int STAGE2_BYTES = A_STAGE_BYTES + B2_STAGE_BYTES;  // 32 KB per CTA per stage
constexpr int SMEM2_BYTES = kStages2 * STAGE2_BYTES + 1024 + 512 + EPI_STAGE_BYTES;

template <int A_MN, int B_MN, int TF32, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
             const GemmParams p) {
    constexpr int kElemBytes = TF32 ? 4 : 2;  // operand element size is a property of the instantiation

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages2 * STAGE2_BYTES);
    uint64_t* empty_bar = full_bar + kStages2;
    uint64_t* tmem_full = empty_bar + kStages2;
    uint64_t* tmem_empty = tmem_full + kAccStages;
    uint64_t* sched_full = tmem_empty + kAccStages;   // [kSched] tile index published (one arrive per CTA)
    uint64_t* sched_empty = sched_full + kSched;      // [kSched] leader's: all 13 consumers of the pair have read it
    uint32_t* sched_tile = reinterpret_cast<uint32_t*>(sched_empty + kSched);  // [kSched]
    uint32_t* tmem_base_slot = sched_tile + kSched;
    float* epi_stage = reinterpret_cast<float*>(smem + kStages2 * STAGE2_BYTES + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_clusters = gridDim.x >> 1;
    const int cluster_id = blockIdx.x >> 1;
    // p.num_m_blocks counts 256-row blocks here
    const int num_tiles = p.num_m_blocks * p.num_n_blocks;
    const int elems_per_row = BLOCK_KB / kElemBytes;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages2; ++i) {
            mbar_init(&full_bar[i], 1);   // leader: one arrive.expect_tx covering both CTAs' bytes
            mbar_init(&empty_bar[i], 1);  // one multicast tcgen05.commit arrive per CTA
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);  // 4 epilogue warps x 2 CTAs arrive on the leader's barrier
        }
        for (int i = 0; i < kSched; ++i) {
            mbar_init(&sched_full[i], 1);
            mbar_init(&sched_empty[i], 13);  // leader: A/B producers, MMA, 4 epilogue warps; peer: the same minus MMA
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == 2) {
        tmem_alloc_2sm(tmem_base_slot, kAccStages * BLOCK_N);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    // ---- dynamic tile scheduler ------------------------------------------------------------------------------
    // Tiles are handed out by an atomic counter (one fetch per cluster per tile) instead of a static stride, so a
    // cluster that cannot become resident (SMs taken by a concurrent NCCL kernel) does not sit on a share of the
    // tiles. The leader's warp 2 publishes each index into BOTH CTAs' shared memory; every consumer role reads it,
    // then arrives on the leader's sched_empty barrier. debug bit 32 = static stride (the old behaviour).
    const bool dyn = !(p.debug & (32 | 16));
    const uint32_t leader_sched_empty0 = mapa_shared(smem_u32(&sched_empty[0]), 0);
    int s_slot = 0, s_static_tile = cluster_id;
    uint32_t s_phase = 0;
    auto next_tile = [&](bool whole_warp) -> int {
        if (!dyn) {
            const int t = s_static_tile;
            s_static_tile += num_clusters;
            return t < num_tiles ? t : -1;
        }
        mbar_wait_cluster(&sched_full[s_slot], s_phase);
        const int t = (int)*reinterpret_cast<volatile uint32_t*>(&sched_tile[s_slot]);
        if (whole_warp) __syncwarp();
        if (!whole_warp || lane == 0) mbar_arrive_remote(leader_sched_empty0 + (uint32_t)s_slot * 8u);
        if (++s_slot == kSched) {
            s_slot = 0;
            s_phase ^= 1;
        }
        return t;
    };
    if (dyn && warp == 2 && lane == 0 && leader) {
        // ===================== tile scheduler (leader CTA) =====================
        int slot = 0;
        uint32_t phase = 0;
        const uint32_t peer_full0 = mapa_shared(smem_u32(&sched_full[0]), 1);
        const uint32_t peer_tile0 = mapa_shared(smem_u32(&sched_tile[0]), 1);
        for (;;) {
            mbar_wait_cluster(&sched_empty[slot], phase ^ 1);
            int t = atomicAdd(p.tile_counter, 1);
            if (t == num_tiles + num_clusters - 1) *p.tile_counter = 0;  // the very last fetch of this launch
            if (t >= num_tiles) t = -1;
            *reinterpret_cast<volatile uint32_t*>(&sched_tile[slot]) = (uint32_t)t;
            st_shared_cluster_u32(peer_tile0 + (uint32_t)slot * 4u, (uint32_t)t);
            mbar_arrive_remote(mapa_shared(smem_u32(&sched_full[slot]), 0));  // release.cluster, own CTA
            mbar_arrive_remote(peer_full0 + (uint32_t)slot * 8u);
            if (t < 0) break;
            if (++slot == kSched) {
                slot = 0;
                phase ^= 1;
            }
        }
    }
    // debug bit 16 (perf experiment): ONE producer thread issues both operands; default: warp 0 loads A and
    // warp 3 loads B so the two TMA instructions of a stage are issued concurrently.
    const bool split_prod = !(p.debug & 16);
    if ((warp == 0 || (warp == 3 && split_prod)) && lane == 0) {
        // ===================== TMA producers (both CTAs) =====================
        const bool do_a = warp == 0, do_b = split_prod ? (warp == 3) : true;
        // Everything that does not change per stage is computed out here: TMA issue from these two threads paces the
        // MN-major kinds (profiles/r1_gemm_mn_major_investigation.md), and a build whose stage loop carried two integer
        // divisions and a mapa lost 11 % on dgrad / wgrad (1403 -> 1248 TFLOP/s, same box, profiles/r1_session2_summary.md).
        const int mn_per_box = 128 / kElemBytes;
        const int box_bytes = elems_per_row * 128;
        const uint32_t bar0 = mapa_shared(smem_u32(&full_bar[0]), 0);  // the leader's full barriers
        const bool arm = leader && do_a;                               // this thread posts the expected byte count
        const int a_mode = A_MN ? (p.a_tma4 == 2 ? 3 : (p.a_tma4 ? 4 : 2)) : ((p.debug & 1) ? 1 : 0);
        const int b_mode = B_MN ? (p.b_tma4 == 2 ? 3 : (p.b_tma4 ? 4 : 2)) : ((p.debug & 1) ? 1 : 0);
        const bool no_tma = (p.debug & 2) != 0;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = next_tile(false); tile >= 0; tile = next_tile(false)) {
            const TileCoord tc = tile_coord(tile, p.num_m_blocks, p.num_n_blocks, p.group_m);
            const int m0 = tc.m_blk * 256 + (int)rank * 128;
            const int n0 = tc.n_blk * BLOCK_N + (int)rank * 128;
            const int m_chunk = m0 / mn_per_box, n_chunk = n0 / mn_per_box;  // once per tile, not per stage
            int k0 = 0;
            for (int kb = 0; kb < p.num_k_blocks; ++kb, k0 += elems_per_row) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE2_BYTES;
                uint8_t* sb = sa + A_STAGE_BYTES;
                const uint32_t bar = bar0 + (uint32_t)stage * 8u;
                if (no_tma) {
                    if (arm) mbar_arrive(&full_bar[stage]);
                } else {
                    if (arm) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);
                    if (do_a) {
                        if (a_mode == 0) {
                            tma_load_2d_2sm(sa, &tma_a, bar, k0, m0);
                        } else if (a_mode == 3) {
                            tma_load_3d_2sm(sa, &tma_a, bar, 0, k0, m_chunk);
                        } else if (a_mode == 4) {
                            tma_load_4d_2sm(sa, &tma_a, bar, 0, 0, m_chunk, k0 >> 3);
                        } else if (a_mode == 1) {
                            tma_load_2d_2sm(sa, &tma_a, bar, k0, m0);
                            tma_load_2d_2sm(sa + 64 * 128, &tma_a, bar, k0, m0 + 64);
                        } else {
                            for (int j = 0; j < 128 / mn_per_box; ++j)
                                tma_load_2d_2sm(sa + j * box_bytes, &tma_a, bar, m0 + j * mn_per_box, k0);
                        }
                    }
                    if (do_b) {
                        if (b_mode == 0) {
                            tma_load_2d_2sm(sb, &tma_b, bar, k0, n0);
                        } else if (b_mode == 3) {
                            tma_load_3d_2sm(sb, &tma_b, bar, 0, k0, n_chunk);
                        } else if (b_mode == 4) {
                            tma_load_4d_2sm(sb, &tma_b, bar, 0, 0, n_chunk, k0 >> 3);
                        } else if (b_mode == 1) {
                            tma_load_2d_2sm(sb, &tma_b, bar, k0, n0);
                            tma_load_2d_2sm(sb + 64 * 128, &tma_b, bar, k0, n0 + 64);
                        } else {
                            for (int j = 0; j < 128 / mn_per_box; ++j)
                                tma_load_2d_2sm(sb + j * box_bytes, &tma_b, bar, n0 + j * mn_per_box, k0);
                        }
                    }
                }
                if (++stage == kStages2) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0 && leader) {
        // ===================== MMA issuer (leader CTA only) =====================
        const uint32_t fmt = TF32 ? 2u : 1u;
        // debug bit 4 (timing experiment only, wrong results): read the operands as MN-major although the TMA
        // loaded them K-major -- separates "MN-major UMMA smem reads" from "MN-major global access pattern".
        const uint32_t amn = (p.debug & 4) ? 1u : ((p.debug & 8) ? 0u : (uint32_t)A_MN);
        const uint32_t bmn = (p.debug & 4) ? 1u : ((p.debug & 8) ? 0u : (uint32_t)B_MN);
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (amn << 15) | (bmn << 16) |
                               ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t mn_box_bytes = elems_per_row * 128;
        // MN-major smem layouts. 2-D path: [chunk][K rows][128 B]   -> LBO = box bytes, SBO = 1024, K-step = rows*128
        //                        4-D path: [K group][chunk][8][128 B] -> LBO = 1024, SBO = chunks*1024 (2 chunks = 128 MN)
        const uint32_t k_rows_per_mma = 32 / kElemBytes;
        const bool a4 = p.a_tma4 == 1, b4 = p.b_tma4 == 1;  // (== 2: 3-D box, same smem layout as the 2-D path)
        const uint32_t a_lbo = amn ? (a4 ? 1024u : mn_box_bytes) : 0u;
        const uint32_t b_lbo = bmn ? (b4 ? 1024u : mn_box_bytes) : 0u;
        const uint32_t a_sbo = (amn && a4) ? 2048u : 1024u;
        const uint32_t b_sbo = (bmn && b4) ? 2048u : 1024u;
        const uint32_t a_kstep = amn ? ((a4 ? (k_rows_per_mma / 8) * 2048u : k_rows_per_mma * 128u) >> 4) : (32 >> 4);
        const uint32_t b_kstep = bmn ? ((b4 ? (k_rows_per_mma / 8) * 2048u : k_rows_per_mma * 128u) >> 4) : (32 >> 4);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = next_tile(false); tile >= 0; tile = next_tile(false), ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BLOCK_N;
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * STAGE2_BYTES);
                const uint32_t sb = sa + A_STAGE_BYTES;
                const uint64_t adesc = make_smem_desc(sa, a_lbo, a_sbo);
                const uint64_t bdesc = make_smem_desc(sb, b_lbo, b_sbo);
#pragma unroll
                for (int k = 0; k < BLOCK_KB / 32; ++k) {
                    const uint64_t ad = adesc + (uint64_t)(k * a_kstep);
                    const uint64_t bd = bdesc + (uint64_t)(k * b_kstep);
                    if (TF32)
                        umma_tf32_2sm(tmem_d, ad, bd, idesc, (kb | k) != 0);
                    else
                        umma_f16_2sm(tmem_d, ad, bd, idesc, (kb | k) != 0);
                }
                umma_commit_2sm(&empty_bar[stage], 3);  // frees the slot in BOTH CTAs
                if (++stage == kStages2) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit_2sm(&tmem_full[as], 3);  // accumulators complete in both CTAs
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own 128 rows) =====================
        const int ew = warp - 4;
        int it = 0;
        const uint32_t leader_empty0 = mapa_shared(smem_u32(&tmem_empty[0]), 0);
        for (int tile = next_tile(true); tile >= 0; tile = next_tile(true), ++it) {
            const TileCoord tc = tile_coord(tile, p.num_m_blocks, p.num_n_blocks, p.group_m);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const long long row0 = (long long)tc.m_blk * 256 + (long long)rank * 128 + ew * 32;
            epilogue_dispatch<EPI>(p, row0, lane, tc.n_blk * BLOCK_N, tmem_base + ((uint32_t)(ew * 32) << 16) + as * BLOCK_N,
                                   epi_stage + ew * (EPI_WARP_BYTES / 4));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_empty0 + (uint32_t)as * 8u);
        }
    }

    tc_fence_before();
    __syncthreads();
    __syncwarp();        // reconverge the single-lane role warps before the warp-aligned cluster barrier
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while the pair still uses its smem / barriers
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, kAccStages * BLOCK_N);
    }
}

// ---------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;
static std::mutex g_mu;
thread_local int g_last_cuda_error = 0;

static int ensure_driver() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_encode) return LLARK_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_EDRIVER;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    return LLARK_OK;
}

struct MapKey {
    const void* ptr;
    long long inner, outer, ld;
    int box_inner, box_outer, elem;
    bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
        return (size_t)h;
    }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

// 2-D tensor map over a row-major [outer][ld] array of which [outer][inner] is valid; 128 B swizzle.
static int get_tensor_map(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long ld,
                          int box_inner, int box_outer, int elem_bytes) {
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = ptr; key.inner = inner; key.outer = outer; key.ld = ld;
    key.box_inner = box_inner; key.box_outer = box_outer; key.elem = elem_bytes;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto itf = g_map_cache.find(key);
        if (itf != g_map_cache.end()) {
            *out = itf->second;
            return LLARK_OK;
        }
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ((ld * elem_bytes) & 15) != 0) return LLARK_EINVAL;
    cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t gstride[1] = {(cuuint64_t)(ld * elem_bytes)};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                          const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = (int)r;
        return LLARK_EDRIVER;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_map_cache.size() > 8192) g_map_cache.clear();
    g_map_cache[key] = *out;
    return LLARK_OK;
}

// Exported to the other translation units of the library (attention_tc.cu): cached 2-D tiled map, 128 B swizzle.
int tensor_map_2d(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long ld, int box_inner,
                  int box_outer, int elem_bytes) {
    int rc = ensure_driver();
    if (rc) return rc;
    return get_tensor_map(out, ptr, inner, outer, ld, box_inner, box_outer, elem_bytes);
}

// 4-D tensor map over a row-major [K][ld] array ([K][MN] valid, MN contiguous) that delivers a whole MN-major
// operand stage with ONE box: dims (128 B of MN, 8 K rows, MN chunks, K groups of 8), box (.., 8, chunks, kgroups).
static int get_tensor_map_mn4(CUtensorMap* out, const void* ptr, long long mn, long long k, long long ld, int chunks,
                              int kgroups, int elem_bytes) {
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = ptr; key.inner = mn; key.outer = k; key.ld = ld;
    key.box_inner = -chunks; key.box_outer = kgroups; key.elem = elem_bytes;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto itf = g_map_cache.find(key);
        if (itf != g_map_cache.end()) {
            *out = itf->second;
            return LLARK_OK;
        }
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ((ld * elem_bytes) & 15) != 0) return LLARK_EINVAL;
    const int epc = 128 / elem_bytes;
    cuuint64_t gdim[4] = {(cuuint64_t)epc, 8, (cuuint64_t)(mn / epc), (cuuint64_t)(k / 8)};
    cuuint64_t gstride[3] = {(cuuint64_t)(ld * elem_bytes), 128, (cuuint64_t)(8 * ld * elem_bytes)};
    cuuint32_t box[4] = {(cuuint32_t)epc, 8, (cuuint32_t)chunks, (cuuint32_t)kgroups};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                          const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = (int)r;
        return LLARK_EDRIVER;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_map_cache.size() > 8192) g_map_cache.clear();  // bounded: variable sequence lengths create new keys
    g_map_cache[key] = *out;
    return LLARK_OK;
}

// 3-D variant: dims (128 B of MN, K rows, MN chunks), box (.., krows, chunks): smem layout [chunk][K rows][128 B].
static int get_tensor_map_mn3(CUtensorMap* out, const void* ptr, long long mn, long long k, long long ld, int chunks,
                              int krows, int elem_bytes) {
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = ptr; key.inner = mn; key.outer = k; key.ld = ld;
    key.box_inner = -100 - chunks; key.box_outer = krows; key.elem = elem_bytes;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto itf = g_map_cache.find(key);
        if (itf != g_map_cache.end()) {
            *out = itf->second;
            return LLARK_OK;
        }
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ((ld * elem_bytes) & 15) != 0) return LLARK_EINVAL;
    const int epc = 128 / elem_bytes;
    cuuint64_t gdim[3] = {(cuuint64_t)epc, (cuuint64_t)k, (cuuint64_t)(mn / epc)};
    cuuint64_t gstride[2] = {(cuuint64_t)(ld * elem_bytes), 128};
    cuuint32_t box[3] = {(cuuint32_t)epc, (cuuint32_t)krows, (cuuint32_t)chunks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                          const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = (int)r;
        return LLARK_EDRIVER;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_map_cache.size() > 8192) g_map_cache.clear();  // bounded: variable sequence lengths create new keys
    g_map_cache[key] = *out;
    return LLARK_OK;
}

template <int A_MN, int B_MN, int TF32, int EPI = 0>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int max_ctas, cudaStream_t stream) {
    static bool attr_set = false;
    auto kfn = gemm_kernel<A_MN, B_MN, TF32, EPI>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) {
            g_last_cuda_error = (int)e;
            return LLARK_ECUDA;
        }
        attr_set = true;
    }
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    int grid = tiles < g_num_sms ? tiles : g_num_sms;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    kfn<<<grid, kThreads, SMEM_BYTES, stream>>>(ta, tb, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

template <int A_MN, int B_MN, int TF32, int EPI = 0>
static int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int max_ctas, cudaStream_t stream) {
    static bool attr_set = false;
    auto kfn = gemm2_kernel<A_MN, B_MN, TF32, EPI>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
        if (e != cudaSuccess) {
            g_last_cuda_error = (int)e;
            return LLARK_ECUDA;
        }
        attr_set = true;
    }
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    int clusters = g_num_sms / 2;
    if (tiles < clusters) clusters = tiles;
    if (max_ctas > 1 && clusters > max_ctas / 2) clusters = max_ctas / 2;
    if (clusters < 1) clusters = 1;
    kfn<<<clusters * 2, kThreads, SMEM2_BYTES, stream>>>(ta, tb, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

static int* g_tile_counters = nullptr;  // ring of zero-initialised device counters for the dynamic scheduler
static unsigned g_counter_seq = 0;
static int* next_tile_counter() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_tile_counters) {
        if (cudaMalloc(&g_tile_counters, 4096 * sizeof(int)) != cudaSuccess) return nullptr;
        cudaMemset(g_tile_counters, 0, 4096 * sizeof(int));
        cudaDeviceSynchronize();  // one-time: the zeroed ring must be visible to launches on any stream
    }
    return g_tile_counters + (g_counter_seq++ % 4096);
}

static int g_gemm_mode = -1;  // -1: read LLARK_GEMM_2CTA from the environment on first use; 0: 1-CTA; 1: CTA pair

}  // namespace llark

using namespace llark;

extern "C" int llark_last_cuda_error(void) { return g_last_cuda_error; }

// One-time, per-process set-up for the current device: resolves the driver entry point used to encode TMA descriptors
// and allocates + zeroes the 16 KB ring of tile counters of the dynamic GEMM scheduler. This is the ONLY place that
// allocates or synchronises; after it every llark_* call only enqueues work on the caller's stream. (llark_gemm still
// initialises lazily if the caller skipped this, so old callers keep working, at the price of a first-call sync.)
extern "C" int llark_init(void) {
    int rc = ensure_driver();
    if (rc != LLARK_OK) return rc;
    return next_tile_counter() ? LLARK_OK : LLARK_ECUDA;
}

extern "C" void llark_gemm_set_mode(int two_cta) { g_gemm_mode = two_cta ? 1 : 0; }
extern "C" int llark_gemm_get_mode(void) {
    if (g_gemm_mode < 0) {
        const char* e = getenv("LLARK_GEMM_2CTA");
        g_gemm_mode = (e && e[0] == '0') ? 0 : 1;
    }
    return g_gemm_mode;
}

extern "C" int llark_gemm(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb,
                          void* C, long long ldc, int c_fp32, const void* bias, int bias_fp32, const void* R,
                          long long ldr, int M, int N, int K, int in_fp32, int max_ctas, cudaStream_t stream) {
    return llark_gemm_ex(A, a_mn_major, lda, B, b_mn_major, ldb, C, ldc, c_fp32, bias, bias_fp32, R, ldr, M, N, K,
                         in_fp32, max_ctas, nullptr, stream);
}

extern "C" int llark_gemm_ex(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                             long long ldb, void* C, long long ldc, int c_fp32, const void* bias, int bias_fp32,
                             const void* R, long long ldr, int M, int N, int K, int in_fp32, int max_ctas,
                             const LlarkGemmEpilogue* ep, cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C) return LLARK_EINVAL;
    if (ep && ep->kind != LLARK_EPI_NONE) {
        switch (ep->kind) {
            case LLARK_EPI_QUICK_GELU:
                if (R) return LLARK_EINVAL;
                break;
            case LLARK_EPI_ROPE:
                // bf16 output, no bias / residual, whole heads per tile half, rotated region ends on a tile boundary
                if (c_fp32 || bias || R || !ep->a || !ep->b || ep->i0 <= 0 || (N % 128) || (ep->i1 % BLOCK_N) ||
                    ep->i1 > N || ((reinterpret_cast<uintptr_t>(ep->a) | reinterpret_cast<uintptr_t>(ep->b)) & 15))
                    return LLARK_EINVAL;
                break;
            case LLARK_EPI_SWIGLU_BWD:
                // C is dgu [M, >= 2N] bf16, ep->a is gu with the same halves
                if (c_fp32 || bias || R || !ep->a || (N % 8) || (ep->ld % 8) || ep->ld < 2LL * N || ldc < 2LL * N ||
                    (reinterpret_cast<uintptr_t>(ep->a) & 15))
                    return LLARK_EINVAL;
                break;
            default:
                return LLARK_EINVAL;
        }
    }
    int rc = ensure_driver();
    if (rc != LLARK_OK) return rc;
    // TF32 operands are supported K-major only: MN-major 32-bit operands need the SWIZZLE_128B_BASE32B atom
    // (measured wrong with the plain 128B swizzle). The Jukebox prior keeps its frozen weights pre-transposed.
    if (in_fp32 && (a_mn_major || b_mn_major)) return LLARK_EINVAL;
    const int eb = in_fp32 ? 4 : 2;
    const int epr = BLOCK_KB / eb;  // K elements per smem row
    const int vec = c_fp32 ? 4 : 8;
    if ((ldc % vec) != 0 || (reinterpret_cast<uintptr_t>(C) & 15) != 0) return LLARK_EINVAL;
    if (R && ((ldr % vec) != 0 || (reinterpret_cast<uintptr_t>(R) & 15) != 0)) return LLARK_EINVAL;

    static int dbg_split = -1;
    if (dbg_split < 0) {
        const char* e = getenv("LLARK_GEMM_DEBUG");
        dbg_split = (e && (atoi(e) & 1)) ? 1 : 0;
    }
    // MN-major operands of the CTA-pair kernel use the one-box 4-D load when the extents tile exactly
    static int use4 = -1;
    if (use4 < 0) {
        const char* e = getenv("LLARK_GEMM_TMA4");
        use4 = e ? atoi(e) : 2;  // 0: 2-D boxes, 1: one 4-D box, 2: one 3-D box (default)
    }
    const int epc_ = 128 / eb;
    const bool a4 = use4 && llark_gemm_get_mode() && a_mn_major && !in_fp32 && (M % epc_ == 0) && (K % 8 == 0);
    const bool b4 = use4 && llark_gemm_get_mode() && b_mn_major && !in_fp32 && (N % epc_ == 0) && (K % 8 == 0);
    CUtensorMap ta, tb;
    if (a4 && use4 == 2)
        rc = get_tensor_map_mn3(&ta, A, M, K, lda, 2, epr, eb);
    else if (a4)
        rc = get_tensor_map_mn4(&ta, A, M, K, lda, 2, epr / 8, eb);
    else if (!a_mn_major)
        rc = get_tensor_map(&ta, A, K, M, lda, epr, dbg_split ? 64 : BLOCK_M, eb);  // [M][K], box {128 B, 128 rows}
    else
        rc = get_tensor_map(&ta, A, M, K, lda, epr, epr, eb);  // [K][M], box {128 B of M, epr K-rows}
    if (rc != LLARK_OK) return rc;
    if (b4 && use4 == 2)
        rc = get_tensor_map_mn3(&tb, B, N, K, ldb, 2, epr, eb);
    else if (b4)
        rc = get_tensor_map_mn4(&tb, B, N, K, ldb, 2, epr / 8, eb);
    else if (!b_mn_major)
        rc = get_tensor_map(&tb, B, K, N, ldb, epr, dbg_split ? 64 : (llark_gemm_get_mode() ? 128 : BLOCK_N), eb);
    else
        rc = get_tensor_map(&tb, B, N, K, ldb, epr, epr, eb);
    if (rc != LLARK_OK) return rc;

    const int two_cta = llark_gemm_get_mode();
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.C = C; p.ldc = ldc; p.c_fp32 = c_fp32;
    p.bias = bias; p.bias_fp32 = bias_fp32;
    p.R = R; p.ldr = ldr;
    p.num_m_blocks = two_cta ? (M + 255) / 256 : (M + BLOCK_M - 1) / BLOCK_M;
    p.num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
    p.num_k_blocks = (K + epr - 1) / epr;
    p.elem_bytes = eb;
    p.tile_counter = nullptr;
    if (two_cta) {
        p.tile_counter = next_tile_counter();
        if (!p.tile_counter) return LLARK_ECUDA;
    }
    p.a_tma4 = a4 ? use4 : 0;
    p.b_tma4 = b4 ? use4 : 0;
    p.epi = ep ? ep->kind : 0;
    {
        static int mode = -1;  // LLARK_GEMM_EPILOGUE = auto (default) | staged | direct
        if (mode < 0) {
            const char* e = getenv("LLARK_GEMM_EPILOGUE");
            mode = (e && e[0] == 's') ? 1 : ((e && e[0] == 'd') ? 2 : 0);
        }
        // auto: staged where the isolated-GEMM probe (tools/gemm_epilogue_probe.py, profiles/r2k_epilogue_probe.md) and
        // the in-step A/B (profiles/r2c_epilogue_ab.md) measured it faster -- the SwiGLU backward; any residual with an fp32
        // tile (K = 1200: 17.8 vs 21.6 us per tile, K = 4800: 35.8 vs 36.2) or with a short K (bf16, K = 1200: 12.8 vs 16.5);
        // fp32 tiles without a residual only when K is short (K = 4800: direct 32.4 vs staged 35.1). Otherwise direct.
        const bool short_k = (long long)K * eb < 2400LL * 2;  // fewer than ~2400 bf16-equivalent K elements
        const bool want = p.epi == LLARK_EPI_SWIGLU_BWD || (R != nullptr && (c_fp32 || short_k)) ||
                          (c_fp32 && R == nullptr && short_k);
        p.staged = mode == 1 ? 1 : (mode == 2 ? 0 : (want ? 1 : 0));
    }
    {
        // Rasterisation group: tiles sweep n inside groups of `group_m` m-blocks, so a group's A panel
        // (group_m x tile_m rows x K) is re-used across the sweep. ncu shows 3.5-4.4x the algorithmic DRAM bytes on the
        // K >= 11008 shapes with the fixed group of 16 (profiles/gemm_traffic.json); LLARK_GEMM_L2_MB=<MB> sizes the group so
        // that panel fits that many MB of L2 instead. Measured (profiles/r2g_negative_results.md): 32 and 48 MB leave every
        // shape's in-step rate within +-1 % -- DRAM is at ~35 % of its bandwidth on these kernels either way -- so the default
        // stays the fixed 16 (0).
        static int l2_mb = -1;
        if (l2_mb < 0) {
            const char* e = getenv("LLARK_GEMM_L2_MB");
            l2_mb = e ? atoi(e) : 0;
        }
        p.group_m = 16;
        if (l2_mb > 0) {
            const double per_mblock = (double)(two_cta ? 256 : BLOCK_M) * K * eb;
            int g = (int)((double)l2_mb * 1e6 / per_mblock);
            p.group_m = g < 2 ? 2 : (g > 32 ? 32 : g);
        }
    }
    p.epi_a = ep ? ep->a : nullptr;
    p.epi_b = ep ? ep->b : nullptr;
    p.epi_ld = ep ? ep->ld : 0;
    p.epi_i0 = ep ? ep->i0 : 0;
    p.epi_i1 = ep ? ep->i1 : 0;
    {
        static int dbg = -1;
        if (dbg < 0) {
            const char* e = getenv("LLARK_GEMM_DEBUG");
            dbg = e ? atoi(e) : 0;
        }
        p.debug = dbg;
    }

    const int sel = (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
    if (p.epi != LLARK_EPI_NONE) {
        // fused epilogues exist for the operand layouts their call sites use (validated above for everything else)
        if (p.epi == LLARK_EPI_QUICK_GELU && sel == 0) {
            if (two_cta) return in_fp32 ? launch2<0, 0, 1, 1>(ta, tb, p, max_ctas, stream)
                                        : launch2<0, 0, 0, 1>(ta, tb, p, max_ctas, stream);
            return in_fp32 ? launch<0, 0, 1, 1>(ta, tb, p, max_ctas, stream)
                           : launch<0, 0, 0, 1>(ta, tb, p, max_ctas, stream);
        }
        if (p.epi == LLARK_EPI_ROPE && sel == 0 && !in_fp32)
            return two_cta ? launch2<0, 0, 0, 2>(ta, tb, p, max_ctas, stream)
                           : launch<0, 0, 0, 2>(ta, tb, p, max_ctas, stream);
        if (p.epi == LLARK_EPI_SWIGLU_BWD && sel == 1 && !in_fp32)
            return two_cta ? launch2<0, 1, 0, 3>(ta, tb, p, max_ctas, stream)
                           : launch<0, 1, 0, 3>(ta, tb, p, max_ctas, stream);
        return LLARK_EINVAL;
    }
    if (two_cta) {
        if (!in_fp32) {
            switch (sel) {
                case 0: return launch2<0, 0, 0>(ta, tb, p, max_ctas, stream);
                case 1: return launch2<0, 1, 0>(ta, tb, p, max_ctas, stream);
                case 2: return launch2<1, 0, 0>(ta, tb, p, max_ctas, stream);
                default: return launch2<1, 1, 0>(ta, tb, p, max_ctas, stream);
            }
        }
        return launch2<0, 0, 1>(ta, tb, p, max_ctas, stream);
    }
    if (!in_fp32) {
        switch (sel) {
            case 0: return launch<0, 0, 0>(ta, tb, p, max_ctas, stream);
            case 1: return launch<0, 1, 0>(ta, tb, p, max_ctas, stream);
            case 2: return launch<1, 0, 0>(ta, tb, p, max_ctas, stream);
            default: return launch<1, 1, 0>(ta, tb, p, max_ctas, stream);
        }
    } else {
        switch (sel) {
            case 0: return launch<0, 0, 1>(ta, tb, p, max_ctas, stream);
            case 1: return launch<0, 1, 1>(ta, tb, p, max_ctas, stream);
            case 2: return launch<1, 0, 1>(ta, tb, p, max_ctas, stream);
            default: return launch<1, 1, 1>(ta, tb, p, max_ctas, stream);
        }
    }
}

// Tensor maps are cached by (pointer, shape): callers that free and re-allocate buffers at the same address
// with another shape get a different key, so the cache is only ever an optimisation. Exposed for tests.
extern "C" void llark_gemm_clear_cache(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_map_cache.clear();
}
