// Causal self-attention with key-padding mask for the Llama decoder (head_dim 128), forward and backward.
// Replaces the eager/SDPA attention inside transformers' LlamaAttention reached at m2t/models/llamav2.py:224
// and the (declared, never wired) flash-attn-1 path m2t/llava/train/llama_flash_attn_monkey_patch.py:12-94.
//
// Layout: packed projections qkv [B*S, 3*NH*128] bf16 (q heads | k heads | v heads per token row, RoPE already
// applied in place), output O [B*S, NH*128] bf16, LSE [B, NH, S] fp32 (natural log; +inf marks a fully masked row).
// This file: the C-ABI entry points and their dispatch, the FA2-style kernels on mma.sync m16n8k16 bf16 (fp32 accumulate,
// online softmax in fp32, exp2 with pre-scaled log2e) that serve S > 4096 and `LLARK_ATTN_TC=0`, the dO.O row-sum
// prep kernel of the backward pass, and the single-token decode kernel of the KV-cache path. The default path for
// S <= 4096 is the tcgen05 / TMEM / TMA implementation in attention_tc.cu (same contract, same tests).
#include "common.cuh"
#include <stdlib.h>
#include "../../include/llark_b200.h"

namespace llark {

extern thread_local int g_last_cuda_error;
int attn_fwd_tc_launch(const void* qkv, const unsigned char* key_mask, void* out, float* lse, int B, int S, int n_heads,
                       float scale, cudaStream_t st);  // attention_tc.cu
int attn_bwd_tc_launch(const void* qkv, const unsigned char* key_mask, const void* dout, const float* lse,
                       const float* Dsum, void* dqkv, int B, int S, int n_heads, float scale, const void* rope_cos,
                       const void* rope_sin, cudaStream_t st);  // rope_cos != null: inverse RoPE of dq | dk fused
// ALiBi instantiation (attention_tc_alibi.cu) for the MPT variant: same kernels + slope_h * key on the scores
int attn_alibi_fwd_tc_launch(const void* qkv, const unsigned char* key_mask, void* out, float* lse, int B, int S,
                             int n_heads, float scale, const float* slopes, cudaStream_t st);
int attn_alibi_bwd_tc_launch(const void* qkv, const unsigned char* key_mask, const void* dout, const float* lse,
                             const float* Dsum, void* dqkv, int B, int S, int n_heads, float scale, const void* rope_cos,
                             const void* rope_sin, const float* slopes, cudaStream_t st);
// LLARK_ATTN_TC bit 0: tcgen05 forward, bit 1: tcgen05 backward. Default 3 (both, validated on B200: same test
// tolerances as the mma.sync kernels, forward 202 -> 89 us and backward 717 -> 481 us at B 16, S 512, 32 heads);
// 0 selects the mma.sync kernels of this file (also used for S > 4096).
static int attn_tc_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("LLARK_ATTN_TC");
        mode = e ? atoi(e) & 3 : 3;
    }
    return mode;
}

constexpr int HD = 128;        // head dim
constexpr int LDS = HD + 8;    // padded smem row (272 B): conflict-free ldmatrix
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ROWS x 128 bf16 tile, global row stride gstride (elements) -> smem stride LDS; rows >= valid zero-filled.
template <int ROWS>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, long long gstride, int valid) {
    for (int i = threadIdx.x; i < ROWS * 16; i += blockDim.x) {
        const int r = i >> 4, c = (i & 15) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < valid) v = *reinterpret_cast<const uint4*>(g + (long long)r * gstride + c);
        *reinterpret_cast<uint4*>(s + r * LDS + c) = v;
    }
}

// Asynchronous variant (cp.async, 16 B per request, zero-fill for rows >= valid): lets the next K/V tile stream
// into the other shared-memory buffer while the tensor cores work on the current one.
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem, bool valid) {
    const uint32_t n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int ROWS>
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* s, const __nv_bfloat16* g, long long gstride, int valid) {
    for (int i = threadIdx.x; i < ROWS * 16; i += blockDim.x) {
        const int r = i >> 4, c = (i & 15) * 8;
        const bool ok = r < valid;
        cp_async_16(s + r * LDS + c, g + (long long)(ok ? r : 0) * gstride + c, ok);
    }
}

// A-operand fragment address (16 rows x 16 k) for ldmatrix.x4: matrices (r0-7,k0-7) (r8-15,k0-7) (r0-7,k8-15) (r8-15,k8-15)
__device__ __forceinline__ const __nv_bfloat16* a_addr(const __nv_bfloat16* s, int row0, int k0, int lane) {
    return s + (row0 + (lane & 15)) * LDS + k0 + (lane >> 4) * 8;
}
// B operand from a [n][k]-stored tile (k contiguous), two n-blocks of 8: (n0-7,k0-7) (n0-7,k8-15) (n8-15,k0-7) (n8-15,k8-15)
__device__ __forceinline__ const __nv_bfloat16* b_addr(const __nv_bfloat16* s, int n0, int k0, int lane) {
    return s + (n0 + (lane & 7) + (lane >> 4) * 8) * LDS + k0 + ((lane >> 3) & 1) * 8;
}
// B operand from a [k][n]-stored tile (n contiguous) via .trans: (k0-7,n0-7) (k8-15,n0-7) (k0-7,n8-15) (k8-15,n8-15)
__device__ __forceinline__ const __nv_bfloat16* bt_addr(const __nv_bfloat16* s, int k0, int n0, int lane) {
    return s + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + n0 + (lane >> 4) * 8;
}

// =============================================================================================== forward
// grid (ceil(S/64), NH, B), 128 threads: each warp owns 16 query rows; key tiles of 64.
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const unsigned char* __restrict__ mask,
                __nv_bfloat16* __restrict__ O, float* __restrict__ lse_out, int S, int NH, float scale_log2) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
    __nv_bfloat16* sKb = sQ + 64 * LDS;        // two K buffers
    __nv_bfloat16* sVb = sKb + 2 * 64 * LDS;   // two V buffers
    float* sMaskb = reinterpret_cast<float*>(sVb + 2 * 64 * LDS);  // two mask rows

    const int qt = gridDim.x - 1 - blockIdx.x;  // longest rows first
    const int h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = qt * 64;
    const long long rs = 3LL * NH * HD;
    const __nv_bfloat16* qptr = qkv + ((long long)b * S + q0) * rs + (long long)h * HD;
    const __nv_bfloat16* kbase = qkv + ((long long)b * S) * rs + (long long)NH * HD + (long long)h * HD;
    const __nv_bfloat16* vbase = kbase + (long long)NH * HD;

    load_tile<64>(sQ, qptr, rs, min(64, S - q0));
    __syncthreads();
    uint32_t qf[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ldsm_x4(qf[ks], a_addr(sQ, warp * 16, ks * 16, lane));

    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    auto prefetch = [&](int kt) {
        const int k0 = kt * 64, buf = kt & 1;
        load_tile_async<64>(sKb + buf * 64 * LDS, kbase + (long long)k0 * rs, rs, min(64, S - k0));
        load_tile_async<64>(sVb + buf * 64 * LDS, vbase + (long long)k0 * rs, rs, min(64, S - k0));
        if (threadIdx.x < 64) {
            const int kk = k0 + threadIdx.x;
            const bool ok = kk < S && (!mask || mask[(long long)b * S + kk] != 0);
            sMaskb[buf * 64 + threadIdx.x] = ok ? 0.f : -INFINITY;
        }
        cp_async_commit();
    };
    prefetch(0);
    for (int kt = 0; kt <= qt; ++kt) {
        if (kt < qt) {
            prefetch(kt + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const __nv_bfloat16* sK = sKb + (kt & 1) * 64 * LDS;
        const __nv_bfloat16* sV = sVb + (kt & 1) * 64 * LDS;
        const float* sMask = sMaskb + (kt & 1) * 64;

        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[i][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t kb[4];
                ldsm_x4(kb, b_addr(sK, j * 16, ks * 16, lane));
                mma16816(s[2 * j], qf[ks], kb[0], kb[1]);
                mma16816(s[2 * j + 1], qf[ks], kb[2], kb[3]);
            }
        }
        float rmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = nb * 8 + (lane & 3) * 2 + (e & 1);
                const int row = warp * 16 + (lane >> 2) + (e >> 1) * 8;
                float v = s[nb][e] * scale_log2 + sMask[col];
                if (kt == qt && col > row) v = -INFINITY;
                s[nb][e] = v;
                rmax[e >> 1] = fmaxf(rmax[e >> 1], v);
            }
        }
        float alpha[2], m_use[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 1));
            rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 2));
            const float m_new = fmaxf(m_run[r], rmax[r]);
            m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
            alpha[r] = exp2f(m_run[r] - m_use[r]);
            m_run[r] = m_new;
            l_run[r] *= alpha[r];
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = exp2f(s[nb][e] - m_use[e >> 1]);
                s[nb][e] = p;
                l_run[e >> 1] += p;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
            o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
        }
#pragma unroll
        for (int ks2 = 0; ks2 < 4; ++ks2) {
            uint32_t pa[4];
            pa[0] = pack_bf16x2(s[2 * ks2][0], s[2 * ks2][1]);
            pa[1] = pack_bf16x2(s[2 * ks2][2], s[2 * ks2][3]);
            pa[2] = pack_bf16x2(s[2 * ks2 + 1][0], s[2 * ks2 + 1][1]);
            pa[3] = pack_bf16x2(s[2 * ks2 + 1][2], s[2 * ks2 + 1][3]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t vb[4];
                ldsm_x4_t(vb, bt_addr(sV, ks2 * 16, j * 16, lane));
                mma16816(o[2 * j], pa, vb[0], vb[1]);
                mma16816(o[2 * j + 1], pa, vb[2], vb[3]);
            }
        }
        __syncthreads();  // all warps are done with this buffer before the next prefetch overwrites it
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + (lane >> 2) + r * 8;
        if (row >= S) continue;
        const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
        __nv_bfloat16* optr = O + ((long long)b * S + row) * ((long long)NH * HD) + (long long)h * HD;
#pragma unroll
        for (int nb = 0; nb < 16; ++nb) {
            const int col = nb * 8 + (lane & 3) * 2;
            *reinterpret_cast<uint32_t*>(optr + col) = pack_bf16x2(o[nb][2 * r] * inv, o[nb][2 * r + 1] * inv);
        }
        if ((lane & 3) == 0)
            lse_out[((long long)b * NH + h) * S + row] =
                l_run[r] > 0.f ? m_run[r] * LN2 + logf(l_run[r]) : INFINITY;
    }
}

// =============================================================================================== decode (KV cache)
// One new query token per sequence against the cached keys/values of the positions before it (and itself): the
// single-token step of m2t/infer.py's model.generate (transformers' cached LlamaAttention, reached from
// m2t/models/llamav2.py:224 with past_key_values). HBM-bound: per (batch row, head) it streams len x 128 bf16 of K
// and of V once. grid (n_heads, B), 128 threads:
//   phase 1: thread j (stride 128) scores key j: dot(q, K[j]) with q in shared memory, K rows read as 16 B vectors;
//   phase 2: block max / sum of exp over the scores kept in shared memory (len <= 4096);
//   phase 3: thread d owns output channel d: sum_j p[j] * V[j][d] (a warp reads 64 contiguous bytes of a V row).
// q: [B, n_heads*128] (RoPE applied); k_cache / v_cache: [B, s_max, n_heads*128]; lens[b] = number of valid cached
// positions of row b INCLUDING the new token; out [B, n_heads*128].
constexpr int DEC_MAX_LEN = 4096;
__global__ void __launch_bounds__(128)
attn_decode_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                   const __nv_bfloat16* __restrict__ vc, const int* __restrict__ lens, __nv_bfloat16* __restrict__ out,
                   int s_max, int NH, float scale_log2) {
    __shared__ float sq[HD];
    __shared__ float sp[DEC_MAX_LEN];
    __shared__ float red[4];
    const int h = blockIdx.x, b = blockIdx.y;
    const int len = min(lens[b], s_max);
    const long long row = (long long)NH * HD;
    const __nv_bfloat16* kb = kc + (long long)b * s_max * row + (long long)h * HD;
    const __nv_bfloat16* vb = vc + (long long)b * s_max * row + (long long)h * HD;
    sq[threadIdx.x] = __bfloat162float(q[(long long)b * row + (long long)h * HD + threadIdx.x]);
    __syncthreads();
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < len; j += 128) {
        const __nv_bfloat16* kr = kb + (long long)j * row;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < HD; c += 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(kr + c);
            const float2 a = unpack_bf16x2(u.x), b2 = unpack_bf16x2(u.y), c2 = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
            dot = fmaf(sq[c], a.x, dot); dot = fmaf(sq[c + 1], a.y, dot);
            dot = fmaf(sq[c + 2], b2.x, dot); dot = fmaf(sq[c + 3], b2.y, dot);
            dot = fmaf(sq[c + 4], c2.x, dot); dot = fmaf(sq[c + 5], c2.y, dot);
            dot = fmaf(sq[c + 6], d.x, dot); dot = fmaf(sq[c + 7], d.y, dot);
        }
        dot *= scale_log2;
        sp[j] = dot;
        mx = fmaxf(mx, dot);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int j = threadIdx.x; j < len; j += 128) {
        const float p = exp2f(sp[j] - mx);
        sp[j] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = red[0] + red[1] + red[2] + red[3];
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    float acc = 0.f;
    for (int j = 0; j < len; ++j) acc = fmaf(sp[j], __bfloat162float(vb[(long long)j * row + threadIdx.x]), acc);
    out[(long long)b * row + (long long)h * HD + threadIdx.x] = __float2bfloat16_rn(acc * inv);
}

// =============================================================================================== backward
// Dsum[b,h,s] = sum_d dO[b,s,h,d] * O[b,s,h,d]; one warp per (b,s,h).
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ O, const __nv_bfloat16* __restrict__ dO,
                                     float* __restrict__ Dsum, long long B, int S, int NH) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long total = B * S * NH;
    if (w >= total) return;
    const int h = (int)(w % NH);
    const long long bs = w / NH;
    const long long off = bs * ((long long)NH * HD) + (long long)h * HD + lane * 4;
    const uint2 a = *reinterpret_cast<const uint2*>(O + off);
    const uint2 d = *reinterpret_cast<const uint2*>(dO + off);
    const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y);
    float s = a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y;
    s = warp_sum(s);
    if (lane == 0) {
        const long long b = bs / S;
        const int sidx = (int)(bs % S);
        Dsum[(b * NH + h) * S + sidx] = s;
    }
}

// dK, dV: grid (ceil(S/64) key tiles, NH, B), 128 threads; each warp owns 16 keys; query blocks of 32.
__global__ void __launch_bounds__(128)
attn_bwd_dkdv_kernel(const __nv_bfloat16* __restrict__ qkv, const unsigned char* __restrict__ mask,
                     const __nv_bfloat16* __restrict__ dO, const float* __restrict__ lse,
                     const float* __restrict__ Dsum, __nv_bfloat16* __restrict__ dqkv, int S, int NH,
                     float scale, float scale_log2) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn);
    __nv_bfloat16* sV = sK + 64 * LDS;
    __nv_bfloat16* sQb = sV + 64 * LDS;            // two Q buffers (32 rows each)
    __nv_bfloat16* sdOb = sQb + 2 * 32 * LDS;      // two dO buffers
    float* sLseb = reinterpret_cast<float*>(sdOb + 2 * 32 * LDS);  // 2 x 32, log2 domain
    float* sDb = sLseb + 64;

    const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = kt * 64;
    const long long rs = 3LL * NH * HD;
    const long long os = (long long)NH * HD;
    const __nv_bfloat16* qbase = qkv + ((long long)b * S) * rs + (long long)h * HD;
    const __nv_bfloat16* kptr = qbase + (long long)NH * HD + (long long)k0 * rs;
    const __nv_bfloat16* vptr = kptr + (long long)NH * HD;
    const __nv_bfloat16* dobase = dO + ((long long)b * S) * os + (long long)h * HD;

    load_tile<64>(sK, kptr, rs, min(64, S - k0));
    load_tile<64>(sV, vptr, rs, min(64, S - k0));

    // key validity for this thread's two key rows (g, g+8)
    bool kvalid[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int kk = k0 + warp * 16 + (lane >> 2) + r * 8;
        kvalid[r] = kk < S && (!mask || mask[(long long)b * S + kk] != 0);
    }

    float dk[16][4], dv[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) dk[i][e] = dv[i][e] = 0.f;

    auto prefetch = [&](int q0, int buf) {
        load_tile_async<32>(sQb + buf * 32 * LDS, qbase + (long long)q0 * rs, rs, min(32, S - q0));
        load_tile_async<32>(sdOb + buf * 32 * LDS, dobase + (long long)q0 * os, os, min(32, S - q0));
        if (threadIdx.x < 32) {
            const int qq = q0 + threadIdx.x;
            const bool ok = qq < S;
            sLseb[buf * 32 + threadIdx.x] = ok ? lse[((long long)b * NH + h) * S + qq] * LOG2E : INFINITY;
            sDb[buf * 32 + threadIdx.x] = ok ? Dsum[((long long)b * NH + h) * S + qq] : 0.f;
        }
        cp_async_commit();
    };
    __syncthreads();  // K/V tiles (synchronous loads above) visible
    prefetch(k0, 0);
    int it = 0;
    for (int q0 = k0; q0 < S; q0 += 32, ++it) {
        if (q0 + 32 < S) {
            prefetch(q0 + 32, (it + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const __nv_bfloat16* sQ = sQb + (it & 1) * 32 * LDS;
        const __nv_bfloat16* sdO = sdOb + (it & 1) * 32 * LDS;
        const float* sLse = sLseb + (it & 1) * 32;
        const float* sD = sDb + (it & 1) * 32;

        // S^T = K Q^T  (16 keys x 32 queries), dP^T = V dO^T
        float st[4][4], dpt[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) st[i][e] = dpt[i][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t ka[4], va[4];
            ldsm_x4(ka, a_addr(sK, warp * 16, ks * 16, lane));
            ldsm_x4(va, a_addr(sV, warp * 16, ks * 16, lane));
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t qb[4], db[4];
                ldsm_x4(qb, b_addr(sQ, j * 16, ks * 16, lane));
                ldsm_x4(db, b_addr(sdO, j * 16, ks * 16, lane));
                mma16816(st[2 * j], ka, qb[0], qb[1]);
                mma16816(st[2 * j + 1], ka, qb[2], qb[3]);
                mma16816(dpt[2 * j], va, db[0], db[1]);
                mma16816(dpt[2 * j + 1], va, db[2], db[3]);
            }
        }
        // P^T and dS^T
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int qc = nb * 8 + (lane & 3) * 2 + (e & 1);  // query within block
                const int kr = warp * 16 + (lane >> 2) + (e >> 1) * 8;  // key within tile
                const bool allowed = kvalid[e >> 1] && (q0 + qc >= k0 + kr);
                const float p = allowed ? exp2f(st[nb][e] * scale_log2 - sLse[qc]) : 0.f;
                st[nb][e] = p;
                dpt[nb][e] = p * (dpt[nb][e] - sD[qc]);
            }
        }
        // dV += P^T dO ; dK += dS^T Q    (k = 32 queries -> 2 k-steps)
#pragma unroll
        for (int ks2 = 0; ks2 < 2; ++ks2) {
            uint32_t pa[4], sa[4];
            pa[0] = pack_bf16x2(st[2 * ks2][0], st[2 * ks2][1]);
            pa[1] = pack_bf16x2(st[2 * ks2][2], st[2 * ks2][3]);
            pa[2] = pack_bf16x2(st[2 * ks2 + 1][0], st[2 * ks2 + 1][1]);
            pa[3] = pack_bf16x2(st[2 * ks2 + 1][2], st[2 * ks2 + 1][3]);
            sa[0] = pack_bf16x2(dpt[2 * ks2][0], dpt[2 * ks2][1]);
            sa[1] = pack_bf16x2(dpt[2 * ks2][2], dpt[2 * ks2][3]);
            sa[2] = pack_bf16x2(dpt[2 * ks2 + 1][0], dpt[2 * ks2 + 1][1]);
            sa[3] = pack_bf16x2(dpt[2 * ks2 + 1][2], dpt[2 * ks2 + 1][3]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t ob[4], qb[4];
                ldsm_x4_t(ob, bt_addr(sdO, ks2 * 16, j * 16, lane));
                ldsm_x4_t(qb, bt_addr(sQ, ks2 * 16, j * 16, lane));
                mma16816(dv[2 * j], pa, ob[0], ob[1]);
                mma16816(dv[2 * j + 1], pa, ob[2], ob[3]);
                mma16816(dk[2 * j], sa, qb[0], qb[1]);
                mma16816(dk[2 * j + 1], sa, qb[2], qb[3]);
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int kk = k0 + warp * 16 + (lane >> 2) + r * 8;
        if (kk >= S) continue;
        __nv_bfloat16* dkp = dqkv + ((long long)b * S + kk) * rs + (long long)NH * HD + (long long)h * HD;
        __nv_bfloat16* dvp = dkp + (long long)NH * HD;
#pragma unroll
        for (int nb = 0; nb < 16; ++nb) {
            const int col = nb * 8 + (lane & 3) * 2;
            *reinterpret_cast<uint32_t*>(dkp + col) = pack_bf16x2(dk[nb][2 * r] * scale, dk[nb][2 * r + 1] * scale);
            *reinterpret_cast<uint32_t*>(dvp + col) = pack_bf16x2(dv[nb][2 * r], dv[nb][2 * r + 1]);
        }
    }
}

// dQ: grid (ceil(S/64) query tiles, NH, B), 128 threads; each warp owns 16 queries; key tiles of 64.
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const unsigned char* __restrict__ mask,
                   const __nv_bfloat16* __restrict__ dO, const float* __restrict__ lse,
                   const float* __restrict__ Dsum, __nv_bfloat16* __restrict__ dqkv, int S, int NH, float scale,
                   float scale_log2) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
    __nv_bfloat16* sdO = sQ + 64 * LDS;
    __nv_bfloat16* sKb = sdO + 64 * LDS;
    __nv_bfloat16* sVb = sKb + 2 * 64 * LDS;
    float* sMaskb = reinterpret_cast<float*>(sVb + 2 * 64 * LDS);  // 2 x 64: 1 = key allowed

    const int qt = gridDim.x - 1 - blockIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = qt * 64;
    const long long rs = 3LL * NH * HD;
    const long long os = (long long)NH * HD;
    const __nv_bfloat16* qptr = qkv + ((long long)b * S + q0) * rs + (long long)h * HD;
    const __nv_bfloat16* kbase = qkv + ((long long)b * S) * rs + (long long)NH * HD + (long long)h * HD;
    const __nv_bfloat16* vbase = kbase + (long long)NH * HD;
    const __nv_bfloat16* doptr = dO + ((long long)b * S + q0) * os + (long long)h * HD;

    load_tile<64>(sQ, qptr, rs, min(64, S - q0));
    load_tile<64>(sdO, doptr, os, min(64, S - q0));

    float lse2[2], dsum[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + (lane >> 2) + r * 8;
        const bool ok = row < S;
        lse2[r] = ok ? lse[((long long)b * NH + h) * S + row] * LOG2E : INFINITY;
        dsum[r] = ok ? Dsum[((long long)b * NH + h) * S + row] : 0.f;
    }

    float dq[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[i][e] = 0.f;

    auto prefetch = [&](int kt) {
        const int k0 = kt * 64, buf = kt & 1;
        load_tile_async<64>(sKb + buf * 64 * LDS, kbase + (long long)k0 * rs, rs, min(64, S - k0));
        load_tile_async<64>(sVb + buf * 64 * LDS, vbase + (long long)k0 * rs, rs, min(64, S - k0));
        if (threadIdx.x < 64) {
            const int kk = k0 + threadIdx.x;
            sMaskb[buf * 64 + threadIdx.x] = (kk < S && (!mask || mask[(long long)b * S + kk] != 0)) ? 1.f : 0.f;
        }
        cp_async_commit();
    };
    prefetch(0);
    for (int kt = 0; kt <= qt; ++kt) {
        if (kt < qt) {
            prefetch(kt + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const __nv_bfloat16* sK = sKb + (kt & 1) * 64 * LDS;
        const __nv_bfloat16* sV = sVb + (kt & 1) * 64 * LDS;
        const float* sMask = sMaskb + (kt & 1) * 64;

        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[i][e] = dp[i][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t qa[4], da[4];
            ldsm_x4(qa, a_addr(sQ, warp * 16, ks * 16, lane));
            ldsm_x4(da, a_addr(sdO, warp * 16, ks * 16, lane));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t kb[4], vb[4];
                ldsm_x4(kb, b_addr(sK, j * 16, ks * 16, lane));
                ldsm_x4(vb, b_addr(sV, j * 16, ks * 16, lane));
                mma16816(s[2 * j], qa, kb[0], kb[1]);
                mma16816(s[2 * j + 1], qa, kb[2], kb[3]);
                mma16816(dp[2 * j], da, vb[0], vb[1]);
                mma16816(dp[2 * j + 1], da, vb[2], vb[3]);
            }
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = nb * 8 + (lane & 3) * 2 + (e & 1);
                const int row = warp * 16 + (lane >> 2) + (e >> 1) * 8;
                const bool allowed = sMask[col] != 0.f && !(kt == qt && col > row);
                const float p = allowed ? exp2f(s[nb][e] * scale_log2 - lse2[e >> 1]) : 0.f;
                s[nb][e] = p * (dp[nb][e] - dsum[e >> 1]);  // dS
            }
        }
#pragma unroll
        for (int ks2 = 0; ks2 < 4; ++ks2) {
            uint32_t sa[4];
            sa[0] = pack_bf16x2(s[2 * ks2][0], s[2 * ks2][1]);
            sa[1] = pack_bf16x2(s[2 * ks2][2], s[2 * ks2][3]);
            sa[2] = pack_bf16x2(s[2 * ks2 + 1][0], s[2 * ks2 + 1][1]);
            sa[3] = pack_bf16x2(s[2 * ks2 + 1][2], s[2 * ks2 + 1][3]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t kb[4];
                ldsm_x4_t(kb, bt_addr(sK, ks2 * 16, j * 16, lane));
                mma16816(dq[2 * j], sa, kb[0], kb[1]);
                mma16816(dq[2 * j + 1], sa, kb[2], kb[3]);
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + (lane >> 2) + r * 8;
        if (row >= S) continue;
        __nv_bfloat16* dqp = dqkv + ((long long)b * S + row) * rs + (long long)h * HD;
#pragma unroll
        for (int nb = 0; nb < 16; ++nb) {
            const int col = nb * 8 + (lane & 3) * 2;
            *reinterpret_cast<uint32_t*>(dqp + col) = pack_bf16x2(dq[nb][2 * r] * scale, dq[nb][2 * r + 1] * scale);
        }
    }
}

constexpr int FWD_SMEM = 5 * 64 * LDS * 2 + 2 * 64 * 4;
constexpr int DKDV_SMEM = (2 * 64 + 4 * 32) * LDS * 2 + 128 * 4;
constexpr int DQ_SMEM = 6 * 64 * LDS * 2 + 2 * 64 * 4;

static int set_smem(const void* fn, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}
static inline int check_launch2() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

}  // namespace llark

using namespace llark;

extern "C" int llark_attn_decode(const void* q, const void* k_cache, const void* v_cache, const int* lens, void* out,
                                 int B, int s_max, int n_heads, int head_dim, float scale, cudaStream_t st) {
    if (head_dim != HD || B <= 0 || s_max <= 0 || s_max > DEC_MAX_LEN || n_heads <= 0 || !lens) return LLARK_EINVAL;
    dim3 grid(n_heads, B);
    attn_decode_kernel<<<grid, 128, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k_cache,
                                             (const __nv_bfloat16*)v_cache, lens, (__nv_bfloat16*)out, s_max, n_heads,
                                             scale * LOG2E);
    return check_launch2();
}

extern "C" int llark_attn_fwd(const void* qkv, const unsigned char* key_mask, void* out, float* lse, int B, int S,
                              int n_heads, int head_dim, float scale, cudaStream_t st) {
    if (head_dim != HD || B <= 0 || S <= 0 || n_heads <= 0) return LLARK_EINVAL;
    // default: the tcgen05/TMEM kernel (attention_tc.cu) for S <= 4096
    if (attn_tc_mode() & 1) {
        const int rc_tc = attn_fwd_tc_launch(qkv, key_mask, out, lse, B, S, n_heads, scale, st);
        if (rc_tc != LLARK_EINVAL) return rc_tc;
    }
    int rc = set_smem((const void*)attn_fwd_kernel, FWD_SMEM);
    if (rc) return rc;
    dim3 grid((S + 63) / 64, n_heads, B);
    attn_fwd_kernel<<<grid, 128, FWD_SMEM, st>>>((const __nv_bfloat16*)qkv, key_mask, (__nv_bfloat16*)out, lse, S,
                                                  n_heads, scale * LOG2E);
    return check_launch2();
}

extern "C" long long llark_attn_bwd_workspace(int B, int S, int n_heads) {
    return (long long)B * S * n_heads * (long long)sizeof(float);
}

// Backward of the causal attention into dqkv = (dq | dk | dv). rope_cos / rope_sin (bf16 [S, 128], null = none): the
// inverse rotary embedding is applied to the dq and dk heads on the way out -- inside the tcgen05 kernels' epilogues
// (attention_tc.cu: store_acc_64_irope), or by llark_rope(inverse) after the mma.sync kernels (S > 4096) -- so that dqkv
// is the gradient of the PRE-RoPE projection output, bit-identical to llark_attn_bwd followed by llark_rope(inverse=1).
static int attn_bwd_impl(const void* qkv, const unsigned char* key_mask, const void* out, const void* dout,
                         const float* lse, void* dqkv, int B, int S, int n_heads, int head_dim, float scale,
                         const void* rope_cos, const void* rope_sin, void* workspace, cudaStream_t st) {
    if (head_dim != HD || B <= 0 || S <= 0 || n_heads <= 0 || !workspace) return LLARK_EINVAL;
    if ((rope_cos == nullptr) != (rope_sin == nullptr)) return LLARK_EINVAL;
    float* Dsum = (float*)workspace;
    const long long warps = (long long)B * S * n_heads;
    attn_bwd_prep_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)out,
                                                                                (const __nv_bfloat16*)dout, Dsum, B, S,
                                                                                n_heads);
    int rc = check_launch2();
    if (rc) return rc;
    if (attn_tc_mode() & 2) {
        const int rc_tc = attn_bwd_tc_launch(qkv, key_mask, dout, lse, Dsum, dqkv, B, S, n_heads, scale, rope_cos,
                                             rope_sin, st);
        if (rc_tc != LLARK_EINVAL) return rc_tc;
    }
    rc = set_smem((const void*)attn_bwd_dkdv_kernel, DKDV_SMEM);
    if (rc) return rc;
    rc = set_smem((const void*)attn_bwd_dq_kernel, DQ_SMEM);
    if (rc) return rc;
    dim3 grid((S + 63) / 64, n_heads, B);
    attn_bwd_dkdv_kernel<<<grid, 128, DKDV_SMEM, st>>>((const __nv_bfloat16*)qkv, key_mask, (const __nv_bfloat16*)dout,
                                                       lse, Dsum, (__nv_bfloat16*)dqkv, S, n_heads, scale, scale * LOG2E);
    rc = check_launch2();
    if (rc) return rc;
    attn_bwd_dq_kernel<<<grid, 128, DQ_SMEM, st>>>((const __nv_bfloat16*)qkv, key_mask, (const __nv_bfloat16*)dout, lse,
                                                   Dsum, (__nv_bfloat16*)dqkv, S, n_heads, scale, scale * LOG2E);
    rc = check_launch2();
    if (rc || !rope_cos) return rc;
    return llark_rope(dqkv, rope_cos, rope_sin, (long long)B * S, S, n_heads, HD, 3LL * n_heads * HD, 1, st);
}

extern "C" int llark_attn_bwd(const void* qkv, const unsigned char* key_mask, const void* out, const void* dout,
                              const float* lse, void* dqkv, int B, int S, int n_heads, int head_dim, float scale,
                              void* workspace, cudaStream_t st) {
    return attn_bwd_impl(qkv, key_mask, out, dout, lse, dqkv, B, S, n_heads, head_dim, scale, nullptr, nullptr, workspace,
                         st);
}
extern "C" int llark_attn_bwd_rope(const void* qkv, const unsigned char* key_mask, const void* out, const void* dout,
                                   const float* lse, void* dqkv, int B, int S, int n_heads, int head_dim, float scale,
                                   const void* rope_cos, const void* rope_sin, void* workspace, cudaStream_t st) {
    if (!rope_cos || !rope_sin) return LLARK_EINVAL;
    return attn_bwd_impl(qkv, key_mask, out, dout, lse, dqkv, B, S, n_heads, head_dim, scale, rope_cos, rope_sin, workspace,
                         st);
}

// ---------------------------------------------------------------------------------------------- ALiBi (MPT variant)
// Causal attention with key-padding mask and the ALiBi bias of MPT
// (m2t/llava/model/mpt/attention.py:31-83, 458-484): scores = q.k / sqrt(128) + slopes[h] * (key - (S - 1)). Same layouts as
// llark_attn_fwd / llark_attn_bwd (packed qkv without RoPE, LSE, Dsum workspace); tcgen05 kernels only (S <= 4096).
extern "C" int llark_attn_alibi_fwd(const void* qkv, const unsigned char* key_mask, void* out, float* lse,
                                    const float* slopes, int B, int S, int n_heads, int head_dim, float scale,
                                    cudaStream_t st) {
    if (head_dim != HD || B <= 0 || S <= 0 || n_heads <= 0 || !slopes) return LLARK_EINVAL;
    return attn_alibi_fwd_tc_launch(qkv, key_mask, out, lse, B, S, n_heads, scale, slopes, st);
}
extern "C" int llark_attn_alibi_bwd(const void* qkv, const unsigned char* key_mask, const void* out, const void* dout,
                                    const float* lse, void* dqkv, const float* slopes, int B, int S, int n_heads,
                                    int head_dim, float scale, void* workspace, cudaStream_t st) {
    if (head_dim != HD || B <= 0 || S <= 0 || n_heads <= 0 || !workspace || !slopes) return LLARK_EINVAL;
    float* Dsum = (float*)workspace;
    const long long warps = (long long)B * S * n_heads;
    attn_bwd_prep_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)out,
                                                                                (const __nv_bfloat16*)dout, Dsum, B, S,
                                                                                n_heads);
    int rc = check_launch2();
    if (rc) return rc;
    return attn_alibi_bwd_tc_launch(qkv, key_mask, dout, lse, Dsum, dqkv, B, S, n_heads, scale, nullptr, nullptr, slopes,
                                    st);
}
