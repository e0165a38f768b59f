// Causal self-attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), head_dim 128: forward kernel
// first, the two backward kernels further down.
// Replaces the eager/SDPA attention inside transformers' LlamaAttention reached at m2t/models/llamav2.py:224 and its
// autograd backward (the same contract as attention.cu's mma.sync kernels, which remain the path for S > 4096).
//
// FORWARD
// One CTA = one (batch, head, 128-query tile); 192 threads, two CTAs resident per SM (99 KB smem, 256 TMEM columns
// each) so one CTA's softmax overlaps the other's MMAs:
//   warp 0 (one lane) : TMA producer. Q tile once (4 boxes of 64 rows x 128 B), then the K / V tiles of 64 keys in the
//                       order the MMA warp consumes them (K0, K1, V0, K2, V1, ...) through a 3-slot 16 KB ring.
//   warp 1            : TMEM allocator; one lane issues tcgen05.mma:
//                         S[j&1] (128 x 64 fp32, TMEM)  = Q (K-major smem) . K_j^T (K-major smem)       8 x K16
//                         O      (128 x 128 fp32, TMEM) += P_j (K-major smem, bf16) . V_j (MN-major smem) 4 x K16
//                       S is double-buffered so Q.K_{j+1}^T runs while the softmax warps work on S_j.
//   warps 2-5         : softmax, one query row per thread (TMEM lane == row): tcgen05.ld S -> scale, causal and
//                       key-padding mask -> running max / sum (exp2 domain) -> bf16 P written to shared memory in the
//                       128-byte-swizzled K-major layout the UMMA descriptor expects. When a row's running max moved,
//                       the O accumulator is rescaled in TMEM (tcgen05.ld / tcgen05.st) before P.V_j is issued.
//                       Epilogue: O / l -> bf16, LSE (natural log, +inf for a fully masked row).
#include "common.cuh"
#include "../../include/llark_b200.h"
#include <cuda.h>

namespace llark {

extern thread_local int g_last_cuda_error;
int tensor_map_2d(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long ld, int box_inner,
                  int box_outer, int elem_bytes);  // gemm_tcgen05.cu (cached cuTensorMapEncodeTiled, 128 B swizzle)

// This file is also compiled a second time by attention_tc_alibi.cu with LLARK_ATC_ALIBI defined: same kernels with an
// ALiBi bias slope_h * key added to the scaled scores (MPT variant, no RoPE involved). Without the macro the
// preprocessed text -- and therefore the generated code of the Llama path -- is exactly what it was before.
#ifndef LLARK_ATC_NS
#define LLARK_ATC_NS atc
#define LLARK_ATC_FWD_LAUNCH attn_fwd_tc_launch
#define LLARK_ATC_BWD_LAUNCH attn_bwd_tc_launch
#endif
#ifdef LLARK_ATC_ALIBI
#define LLARK_ATC_SLOPES_PARAM , const float* __restrict__ alibi_slopes
#define LLARK_ATC_SLOPES_ARG , alibi_slopes
#else
#define LLARK_ATC_SLOPES_PARAM
#define LLARK_ATC_SLOPES_ARG
#endif

namespace LLARK_ATC_NS {

constexpr int HD = 128;
constexpr int BQ = 128;         // query rows per CTA
constexpr int BK = 64;          // keys per tile
constexpr int kThreads = 192;
constexpr int kSlots = 3;       // K/V ring
constexpr int Q_BYTES = BQ * HD * 2;        // 32 KB: 2 d-chunks x 128 rows x 128 B
constexpr int KV_BYTES = BK * HD * 2;       // 16 KB: 2 d-chunks x 64 rows x 128 B
constexpr int P_BYTES = BQ * BK * 2;        // 16 KB: 128 rows x 128 B
constexpr int MAX_S = 4096;                 // key-mask bit words kept in shared memory
constexpr int OFF_KV = Q_BYTES;
constexpr int OFF_P = OFF_KV + kSlots * KV_BYTES;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int OFF_MASK = OFF_BAR + 128;
constexpr int SMEM_BYTES = OFF_MASK + (MAX_S / 32) * 4 + 1024 /*alignment slack*/;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void softmax_group_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct Ring {
    int slot = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++slot == kSlots) {
            slot = 0;
            phase ^= 1;
        }
    }
};

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const unsigned char* __restrict__ mask,
                   __nv_bfloat16* __restrict__ O, float* __restrict__ lse_out, int S, int NH, float scale_log2
                   LLARK_ATC_SLOPES_PARAM) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + OFF_KV;
    uint8_t* sP = smem + OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars;              // [1]
    uint64_t* kv_full = bars + 1;         // [kSlots]
    uint64_t* kv_empty = kv_full + kSlots;  // [kSlots]
    uint64_t* s_full = kv_empty + kSlots;   // [2]
    uint64_t* s_empty = s_full + 2;         // [2]
    uint64_t* p_full = s_empty + 2;         // [1]
    uint64_t* pv_done = p_full + 1;         // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
    uint32_t* maskw = reinterpret_cast<uint32_t*>(smem + OFF_MASK);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = gridDim.x - 1 - blockIdx.x;  // longest rows first
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * BQ;
    const int n_key_tiles = (S + BK - 1) / BK;
    const int n_kv = min(n_key_tiles, (q0 + BQ) / BK);  // causal: keys <= q0 + 127

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        mbar_init(q_full, 1);
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 4);
        }
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_O = tmem_base + 128;  // columns [128, 256); S buffers at [0, 64) and [64, 128)

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            const int row0 = b * S;
            const int cq = h * HD, ck = NH * HD + h * HD, cv = 2 * NH * HD + h * HD;
            mbar_arrive_expect_tx(q_full, Q_BYTES);
            for (int c = 0; c < 2; ++c)
                for (int rh = 0; rh < 2; ++rh)
                    tma_load_2d(sQ + c * (BQ * 128) + rh * (64 * 128), &tm_qkv, q_full, cq + c * 64, row0 + q0 + rh * 64);
            Ring r;
            auto load_tile = [&](int col, int j) {
                mbar_wait(&kv_empty[r.slot], r.phase ^ 1);
                uint8_t* dst = sKV + r.slot * KV_BYTES;
                mbar_arrive_expect_tx(&kv_full[r.slot], KV_BYTES);
                tma_load_2d(dst, &tm_qkv, &kv_full[r.slot], col, row0 + j * BK);
                tma_load_2d(dst + BK * 128, &tm_qkv, &kv_full[r.slot], col + 64, row0 + j * BK);
                r.advance();
            };
            load_tile(ck, 0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) load_tile(ck, j + 1);
                load_tile(cv, j);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            // instruction descriptors: F32 accumulate, bf16 operands; S: M128 N64, both K-major;
            // PV: M128 N128, A (P) K-major, B (V, stored [key][d]) MN-major
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BK >> 3) << 17) |
                                     ((uint32_t)(BQ >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) |
                                      ((uint32_t)(BQ >> 4) << 24);
            const uint64_t qdesc0 = make_desc(smem_u32(sQ), 0, 1024);
            const uint64_t qdesc1 = make_desc(smem_u32(sQ + BQ * 128), 0, 1024);
            const uint64_t pdesc = make_desc(smem_u32(sP), 0, 1024);
            Ring r;
            auto issue_s = [&](int j) {
                const int sb = j & 1;
                mbar_wait(&s_empty[sb], ((uint32_t)(j >> 1) & 1) ^ 1);
                mbar_wait(&kv_full[r.slot], r.phase);
                tc_fence_after();
                const uint32_t kaddr = smem_u32(sKV + r.slot * KV_BYTES);
                const uint64_t kdesc0 = make_desc(kaddr, 0, 1024);
                const uint64_t kdesc1 = make_desc(kaddr + BK * 128, 0, 1024);
                const uint32_t tS = tmem_base + sb * BK;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t ad = (k < 4 ? qdesc0 : qdesc1) + (uint64_t)((k & 3) * 2);
                    const uint64_t bd = (k < 4 ? kdesc0 : kdesc1) + (uint64_t)((k & 3) * 2);
                    umma_f16(tS, ad, bd, idesc_s, k != 0);
                }
                umma_commit(&kv_empty[r.slot]);
                umma_commit(&s_full[sb]);
                r.advance();
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) issue_s(j + 1);
                mbar_wait(&kv_full[r.slot], r.phase);
                mbar_wait(p_full, (uint32_t)j & 1);
                tc_fence_after();
                const uint64_t vdesc = make_desc(smem_u32(sKV + r.slot * KV_BYTES), BK * 128, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_O, pdesc + (uint64_t)(k * 2), vdesc + (uint64_t)(k * 128), idesc_pv, (j | k) != 0);
                umma_commit(&kv_empty[r.slot]);
                umma_commit(pv_done);
                r.advance();
            }
        }
    } else {
        // ===================== softmax / epilogue (warps 2-5) =====================
        const int quad = warp & 3;             // TMEM lane quadrant this warp may access
        const int r_in = quad * 32 + lane;     // row inside the query tile == TMEM lane
        const int qi = q0 + r_in;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        // key-padding mask -> one bit per key
        {
            const int sw = warp - 2;
            const int nwords = n_key_tiles * 2;
            for (int w = sw; w < nwords; w += 4) {
                const int key = w * 32 + lane;
                const bool ok = key < S && (!mask || mask[(long long)b * S + key] != 0);
                const uint32_t bits = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) maskw[w] = bits;
            }
            softmax_group_sync();
        }
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kv; ++j) {
            const int sb = j & 1;
            mbar_wait(&s_full[sb], (uint32_t)(j >> 1) & 1);
            tc_fence_after();
            uint32_t sv[2][32];
            tmem_ld_32x32b_x32(lane_base + sb * BK, sv[0]);
            tmem_ld_32x32b_x32(lane_base + sb * BK + 32, sv[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);

            const int k0 = j * BK;
            const uint32_t mw0 = maskw[2 * j], mw1 = maskw[2 * j + 1];
            const bool plain = (k0 + BK - 1 <= q0 + quad * 32) && mw0 == 0xffffffffu && mw1 == 0xffffffffu;  // warp-uniform
#ifdef LLARK_ATC_ALIBI
            // ALiBi variant: v = s * scale * log2e + slope_h * log2e * (key - (S - 1)) (the reference's bias, <= 0) is formed first (the bias differs per key, so
            // the row max has to be taken over v, not over the raw scores); masked entries -> -inf
            const float slope2 = alibi_slopes[h] * LOG2E;
            float rmax = -INFINITY;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint32_t mw = hh ? mw1 : mw0;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int key = k0 + hh * 32 + c;
                    const bool ok = plain || (key <= qi && ((mw >> c) & 1u));
                    const float v = ok ? fmaf(__uint_as_float(sv[hh][c]), scale_log2, slope2 * (float)(key - (S - 1))) : -INFINITY;
                    sv[hh][c] = __float_as_uint(v);
                    rmax = fmaxf(rmax, v);
                }
            }
            const float m_cand = fmaxf(m_run, rmax);
            const bool move = (m_cand > m_run + 8.f) || (m_run == -INFINITY);
            const float m_new = move ? m_cand : m_run;
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = ex2(m_run - m_use);
            m_run = m_new;
            float psum = 0.f;
            uint32_t pk[32];  // 64 bf16 probabilities
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float p0 = ex2(__uint_as_float(sv[hh][c]) - m_use);
                    const float p1 = ex2(__uint_as_float(sv[hh][c + 1]) - m_use);
                    psum += p0 + p1;
                    pk[hh * 16 + (c >> 1)] = pack_bf16x2(p0, p1);
                }
#else
            // row max of the RAW scores (scale > 0, so the max commutes with the scaling); masked entries -> -inf
            float rmax = -INFINITY;
            if (plain) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                    for (int c = 0; c < 32; ++c) rmax = fmaxf(rmax, __uint_as_float(sv[hh][c]));
            } else {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const uint32_t mw = hh ? mw1 : mw0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int key = k0 + hh * 32 + c;
                        const bool ok = key <= qi && ((mw >> c) & 1u);
                        const float v = ok ? __uint_as_float(sv[hh][c]) : -INFINITY;
                        sv[hh][c] = __float_as_uint(v);
                        rmax = fmaxf(rmax, v);
                    }
                }
            }
            // Lazy rescaling: the reference maximum m_run only moves when the true maximum has grown by more than
            // 2^8 (probabilities then stay <= 256, exact enough in bf16 / fp32), so the TMEM round trip that rescales
            // O is rare after the first tile. O / l and LSE = m*ln2 + log(l) are invariant to the choice of m_run.
            const float m_cand = fmaxf(m_run, rmax * scale_log2);
            const bool move = (m_cand > m_run + 8.f) || (m_run == -INFINITY);
            const float m_new = move ? m_cand : m_run;
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = ex2(m_run - m_use);
            m_run = m_new;
            float psum = 0.f;
            uint32_t pk[32];  // 64 bf16 probabilities
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float p0 = ex2(fmaf(__uint_as_float(sv[hh][c]), scale_log2, -m_use));
                    const float p1 = ex2(fmaf(__uint_as_float(sv[hh][c + 1]), scale_log2, -m_use));
                    psum += p0 + p1;
                    pk[hh * 16 + (c >> 1)] = pack_bf16x2(p0, p1);
                }
#endif
            l_run = l_run * alpha + psum;

            if (j > 0) {
                // P.V_{j-1} must have retired: sP is free again and O holds tiles 0..j-1
                mbar_wait(pv_done, (uint32_t)(j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
                    for (int c = 0; c < HD / 32; ++c) {
                        uint32_t ov[32];
                        tmem_ld_32x32b_x32(lane_base + 128 + c * 32, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                        tmem_st_32x32b_x32(lane_base + 128 + c * 32, ov);
                    }
                    tmem_st_wait();
                }
            }
            // P row -> shared memory, K-major 128 B rows, 16-byte units XOR-swizzled by (row & 7)
            {
                uint8_t* prow = sP + r_in * 128;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint4 val = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                    *reinterpret_cast<uint4*>(prow + ((u ^ (r_in & 7)) << 4)) = val;
                }
            }
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // epilogue: O / l -> bf16
        mbar_wait(pv_done, (uint32_t)(n_kv - 1) & 1);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        const bool row_ok = qi < S;
        __nv_bfloat16* optr = O + ((long long)b * S + qi) * ((long long)NH * HD) + (long long)h * HD;
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(lane_base + 128 + c * 32, ov);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(ov[8 * u]) * inv, __uint_as_float(ov[8 * u + 1]) * inv);
                    o.y = pack_bf16x2(__uint_as_float(ov[8 * u + 2]) * inv, __uint_as_float(ov[8 * u + 3]) * inv);
                    o.z = pack_bf16x2(__uint_as_float(ov[8 * u + 4]) * inv, __uint_as_float(ov[8 * u + 5]) * inv);
                    o.w = pack_bf16x2(__uint_as_float(ov[8 * u + 6]) * inv, __uint_as_float(ov[8 * u + 7]) * inv);
                    *reinterpret_cast<uint4*>(optr + c * 32 + u * 8) = o;
                }
            }
        }
        if (row_ok)
            lse_out[((long long)b * NH + h) * S + qi] = l_run > 0.f ? m_run * LN2 + logf(l_run) : INFINITY;
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}


// =====================================================================================================================
// BACKWARD on tcgen05. Two kernels with the forward kernel's structure and exactly its two MMA forms
//   "KK": D[128 x 64]  = A[128 rows, K-major, d in 2 chunks] . B[64 rows, K-major]^T            (8 x K16)
//   "PV": D[128 x 128] += A[128 rows x 64, K-major, written by threads] . B[64 K-rows x 128, MN-major]  (4 x K16)
// so no operand is ever transposed in memory:
//   dQ kernel   (CTA = 128 queries, streams 64-key tiles):  S = Q K^T, dP = dO V^T (KK); dS = P o (dP - D);
//                                                           dQ += dS K (PV, K_j read MN-major)
//   dK/dV kernel (CTA = 128 keys, streams 64-query tiles):  S^T = K Q^T, dP^T = V dO^T (KK); P^T, dS^T;
//                                                           dV += P^T dO, dK += dS^T Q (PV, dO_i / Q_i read MN-major)
// P is recomputed from the saved LSE (p = exp2(s*scale*log2e - lse*log2e)), D = rowsum(dO o O) comes from
// attn_bwd_prep_kernel (attention.cu). 320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-9 = two compute warpgroups that split the 64 columns of every S / dP row (no row reductions are needed in
// the backward pass, so the split is free). S and dP are double-buffered in TMEM: the KK MMAs of tile t+1 run while
// the compute warps work on tile t. One CTA per SM (145-162 KB of shared memory, all 512 TMEM columns).
constexpr int kBwdThreads = 320;
constexpr int RING_SLOTS = 6;   // three streamed tiles in flight: with 4 the load of tile j+2 could only be requested after
                                // tile j's last MMA retired, and its latency sat on the critical path of every tile
__device__ __forceinline__ int ring_slot(int it) { return it % RING_SLOTS; }
__device__ __forceinline__ uint32_t ring_phase(int it) { return (uint32_t)(it / RING_SLOTS) & 1u; }

__device__ __forceinline__ void mma_kk_128x64(uint32_t tD, uint32_t a_addr, uint32_t a_chunk_bytes, uint32_t b_addr,
                                              uint32_t b_chunk_bytes, uint32_t idesc) {
    const uint64_t a0 = make_desc(a_addr, 0, 1024), a1 = make_desc(a_addr + a_chunk_bytes, 0, 1024);
    const uint64_t b0 = make_desc(b_addr, 0, 1024), b1 = make_desc(b_addr + b_chunk_bytes, 0, 1024);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        umma_f16(tD, (k < 4 ? a0 : a1) + (uint64_t)((k & 3) * 2), (k < 4 ? b0 : b1) + (uint64_t)((k & 3) * 2), idesc,
                 k != 0);
}
__device__ __forceinline__ void mma_pv_128x128(uint32_t tD, uint32_t a_addr, uint32_t b_addr, uint32_t idesc,
                                               bool accumulate) {
    const uint64_t ad = make_desc(a_addr, 0, 1024);
    const uint64_t bd = make_desc(b_addr, BK * 128, 1024);  // MN-major: the two 64-wide N chunks are 8 KB apart
#pragma unroll
    for (int k = 0; k < 4; ++k)
        umma_f16(tD, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 128), idesc, accumulate || k != 0);
}
__device__ __forceinline__ void compute_group_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// 32 fp32 values -> 32 bf16 -> four 16-byte units (g*4 .. g*4+3) of row r of a 128-byte-row, XOR-swizzled tile
__device__ __forceinline__ void store_half_row(uint8_t* tile, int r, int g, const float (&v)[32]) {
    uint8_t* row = tile + r * 128;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        uint4 w;
        w.x = pack_bf16x2(v[8 * u], v[8 * u + 1]);
        w.y = pack_bf16x2(v[8 * u + 2], v[8 * u + 3]);
        w.z = pack_bf16x2(v[8 * u + 4], v[8 * u + 5]);
        w.w = pack_bf16x2(v[8 * u + 6], v[8 * u + 7]);
        *reinterpret_cast<uint4*>(row + (((g * 4 + u) ^ (r & 7)) << 4)) = w;
    }
}
// 64 fp32 accumulator columns of this thread's TMEM lane -> scaled bf16 -> 128 contiguous bytes in global memory
__device__ __forceinline__ void store_acc_64(uint32_t taddr, __nv_bfloat16* dst, float mul, bool ok) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(taddr + c * 32, ov);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(ov[8 * u]) * mul, __uint_as_float(ov[8 * u + 1]) * mul);
                o.y = pack_bf16x2(__uint_as_float(ov[8 * u + 2]) * mul, __uint_as_float(ov[8 * u + 3]) * mul);
                o.z = pack_bf16x2(__uint_as_float(ov[8 * u + 4]) * mul, __uint_as_float(ov[8 * u + 5]) * mul);
                o.w = pack_bf16x2(__uint_as_float(ov[8 * u + 6]) * mul, __uint_as_float(ov[8 * u + 7]) * mul);
                *reinterpret_cast<uint4*>(dst + c * 32 + u * 8) = o;
            }
        }
    }
}

// The same 64 output columns with the INVERSE rotary embedding applied on the way out (the gradient of
// apply_rotary_pos_emb w.r.t. its input, what llark_rope(inverse=1) does to the dq | dk columns after an unfused
// backward). The rotation partner of column d is column d +- 64 of the same row, i.e. the OTHER half of this thread's
// TMEM lane, so every thread reads both halves and writes its own:
//   half 0:  o[d]      = bf16(x[d] * cos[d])           + bf16( x[d + 64] * sin[d + 64])
//   half 1:  o[d + 64] = bf16(x[d + 64] * cos[d + 64]) + bf16(-x[d] * sin[d])        with x = bf16(acc * mul),
// the rounding points of rope_kernel (llama_ops.cu) on the bf16 values the unfused path would have stored: bit-identical.
// cs / sn: this row's 128 cos / sin values (bf16, position = row % S).
__device__ __forceinline__ void store_acc_64_irope(uint32_t taddr_row, int g, __nv_bfloat16* dst, float mul, bool ok,
                                                   const __nv_bfloat16* cs, const __nv_bfloat16* sn) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        uint32_t own[32], oth[32];
        tmem_ld_32x32b_x32(taddr_row + g * 64 + c * 32, own);
        tmem_ld_32x32b_x32(taddr_row + (1 - g) * 64 + c * 32, oth);
        tmem_ld_wait();
        if (ok) {
            const int d0 = c * 32;  // index inside the half
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                // cos of the OWN column, sin of the PARTNER's column (see the formulas above)
                const uint4 cw = *reinterpret_cast<const uint4*>(cs + g * 64 + d0 + u * 8);
                const uint4 sw = *reinterpret_cast<const uint4*>(sn + (1 - g) * 64 + d0 + u * 8);
                const uint32_t cwv[4] = {cw.x, cw.y, cw.z, cw.w}, swv[4] = {sw.x, sw.y, sw.z, sw.w};
                uint32_t ow[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    const float2 cc = unpack_bf16x2(cwv[e2]), ss = unpack_bf16x2(swv[e2]);
                    float r[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int e = u * 8 + e2 * 2 + q;
                        const float xo = bf16_round(__uint_as_float(own[e]) * mul);
                        const float xp = bf16_round(__uint_as_float(oth[e]) * mul);
                        const float cv = q ? cc.y : cc.x, sv = q ? ss.y : ss.x;
                        r[q] = bf16_round(xo * cv) + bf16_round((g ? -xp : xp) * sv);
                    }
                    ow[e2] = pack_bf16x2(r[0], r[1]);
                }
                *reinterpret_cast<uint4*>(dst + d0 + u * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------- dQ
constexpr int DQ_OFF_DO = Q_BYTES;
constexpr int DQ_OFF_RING = 2 * Q_BYTES;
constexpr int DQ_OFF_DS = DQ_OFF_RING + RING_SLOTS * KV_BYTES;
constexpr int DQ_OFF_BAR = DQ_OFF_DS + 2 * P_BYTES;  // dS is double-buffered
constexpr int DQ_OFF_MASK = DQ_OFF_BAR + 256;
constexpr int DQ_SMEM_BYTES = DQ_OFF_MASK + (MAX_S / 32) * 4 + 1024;

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                      const unsigned char* __restrict__ mask, const float* __restrict__ lse,
                      const float* __restrict__ Dsum, __nv_bfloat16* __restrict__ dqkv, int S, int NH, float scale,
                      float scale_log2, const __nv_bfloat16* __restrict__ rope_cos,
                      const __nv_bfloat16* __restrict__ rope_sin LLARK_ATC_SLOPES_PARAM) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sdO = smem + DQ_OFF_DO;
    uint8_t* sRing = smem + DQ_OFF_RING;
    uint8_t* sdS = smem + DQ_OFF_DS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_OFF_BAR);
    uint64_t* q_full = bars;                       // [1]  Q and dO tiles landed
    uint64_t* ring_full = bars + 1;                // [RING_SLOTS]
    uint64_t* ring_empty = ring_full + RING_SLOTS; // [RING_SLOTS]
    uint64_t* sdp_full = ring_empty + RING_SLOTS;  // [2]  S and dP of tile j are in TMEM
    uint64_t* sdp_empty = sdp_full + 2;            // [2]  8 compute warps have read them
    uint64_t* ds_full = sdp_empty + 2;             // [2]  dS tile j written to shared-memory buffer j & 1
    uint64_t* dq_done = ds_full + 2;               // [2]  dQ += dS_j K_j retired (buffer j & 1 free again)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_done + 2);
    uint32_t* maskw = reinterpret_cast<uint32_t*>(smem + DQ_OFF_MASK);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = gridDim.x - 1 - blockIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * BQ;
    const int n_key_tiles = (S + BK - 1) / BK;
    const int n_kv = min(n_key_tiles, (q0 + BQ) / BK);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        mbar_init(q_full, 1);
        for (int i = 0; i < RING_SLOTS; ++i) {
            mbar_init(&ring_full[i], 1);
            mbar_init(&ring_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sdp_full[i], 1);
            mbar_init(&sdp_empty[i], 8);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&ds_full[i], 8);
            mbar_init(&dq_done[i], 1);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;  // S[2] @ 0,64   dP[2] @ 128,192   dQ @ 256..383

    if (warp == 0) {
        if (lane == 0) {
            const int row0 = b * S;
            const int cq = h * HD, ck = NH * HD + h * HD, cv = 2 * NH * HD + h * HD;
            mbar_arrive_expect_tx(q_full, 2 * Q_BYTES);
            for (int c = 0; c < 2; ++c)
                for (int rh = 0; rh < 2; ++rh) {
                    tma_load_2d(sQ + c * (BQ * 128) + rh * (64 * 128), &tm_qkv, q_full, cq + c * 64, row0 + q0 + rh * 64);
                    tma_load_2d(sdO + c * (BQ * 128) + rh * (64 * 128), &tm_do, q_full, cq + c * 64, row0 + q0 + rh * 64);
                }
            for (int j = 0; j < n_kv; ++j) {
#pragma unroll
                for (int kv = 0; kv < 2; ++kv) {
                    const int it = 2 * j + kv, slot = ring_slot(it);
                    mbar_wait(&ring_empty[slot], ring_phase(it) ^ 1);
                    uint8_t* dst = sRing + slot * KV_BYTES;
                    mbar_arrive_expect_tx(&ring_full[slot], KV_BYTES);
                    const int col = kv ? cv : ck;
                    tma_load_2d(dst, &tm_qkv, &ring_full[slot], col, row0 + j * BK);
                    tma_load_2d(dst + BK * 128, &tm_qkv, &ring_full[slot], col + 64, row0 + j * BK);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_kk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BK >> 3) << 17) |
                                      ((uint32_t)(BQ >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) |
                                      ((uint32_t)(BQ >> 4) << 24);
            auto issue_sdp = [&](int j) {
                const int sb = j & 1;
                const int itk = 2 * j, itv = 2 * j + 1;
                mbar_wait(&sdp_empty[sb], ((uint32_t)(j >> 1) & 1) ^ 1);
                mbar_wait(&ring_full[ring_slot(itk)], ring_phase(itk));
                tc_fence_after();
                mma_kk_128x64(tmem_base + sb * BK, smem_u32(sQ), BQ * 128, smem_u32(sRing + ring_slot(itk) * KV_BYTES),
                              BK * 128, idesc_kk);
                mbar_wait(&ring_full[ring_slot(itv)], ring_phase(itv));
                tc_fence_after();
                mma_kk_128x64(tmem_base + 128 + sb * BK, smem_u32(sdO), BQ * 128,
                              smem_u32(sRing + ring_slot(itv) * KV_BYTES), BK * 128, idesc_kk);
                umma_commit(&ring_empty[ring_slot(itv)]);  // V_j is not needed again
                umma_commit(&sdp_full[sb]);
            };
            mbar_wait(q_full, 0);
            issue_sdp(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) issue_sdp(j + 1);
                mbar_wait(&ds_full[j & 1], (uint32_t)(j >> 1) & 1);
                tc_fence_after();
                const int itk = 2 * j;
                mma_pv_128x128(tmem_base + 256, smem_u32(sdS + (j & 1) * P_BYTES),
                               smem_u32(sRing + ring_slot(itk) * KV_BYTES), idesc_pv, j != 0);
                umma_commit(&ring_empty[ring_slot(itk)]);
                umma_commit(&dq_done[j & 1]);
            }
        }
    } else {
        const int cw = warp - 2;
        const int g = cw >> 2;       // column half of every S / dP row this warpgroup handles
        const int quad = warp & 3;
        const int r_in = quad * 32 + lane;
        const int qi = q0 + r_in;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        {
            const int nwords = n_key_tiles * 2;
            for (int w = cw; w < nwords; w += 8) {
                const int key = w * 32 + lane;
                const bool ok = key < S && (!mask || mask[(long long)b * S + key] != 0);
                const uint32_t bits = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) maskw[w] = bits;
            }
            compute_group_sync();
        }
        const bool row_ok = qi < S;
        const float lse2 = row_ok ? lse[((long long)b * NH + h) * S + qi] * LOG2E : INFINITY;
        const float dsum = row_ok ? Dsum[((long long)b * NH + h) * S + qi] : 0.f;
        for (int j = 0; j < n_kv; ++j) {
            const int sb = j & 1;
            mbar_wait(&sdp_full[sb], (uint32_t)(j >> 1) & 1);
            tc_fence_after();
            uint32_t sv[32], dv[32];
            tmem_ld_32x32b_x32(lane_base + sb * BK + g * 32, sv);
            tmem_ld_32x32b_x32(lane_base + 128 + sb * BK + g * 32, dv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sdp_empty[sb]);

            const int kbase = j * BK + g * 32;
            const uint32_t mw = maskw[2 * j + g];
            const bool plain = (kbase + 31 <= q0 + quad * 32) && mw == 0xffffffffu;  // warp-uniform
            float ds[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const bool ok = plain || (kbase + c <= qi && ((mw >> c) & 1u));
#ifdef LLARK_ATC_ALIBI
                const float p = ok ? ex2(fmaf(__uint_as_float(sv[c]), scale_log2,
                                             fmaf(alibi_slopes[h] * LOG2E, (float)(kbase + c - (S - 1)), -lse2))) : 0.f;
#else
                const float p = ok ? ex2(fmaf(__uint_as_float(sv[c]), scale_log2, -lse2)) : 0.f;
#endif
                ds[c] = p * (__uint_as_float(dv[c]) - dsum);
            }
            if (j >= 2) {
                mbar_wait(&dq_done[j & 1], (uint32_t)((j >> 1) - 1) & 1);  // tile j-2 retired: buffer j & 1 is free
                tc_fence_after();
            }
            store_half_row(sdS + (j & 1) * P_BYTES, r_in, g, ds);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ds_full[j & 1]);
        }
        mbar_wait(&dq_done[(n_kv - 1) & 1], (uint32_t)((n_kv - 1) >> 1) & 1);  // MMAs retire in order
        tc_fence_after();
        __nv_bfloat16* dqp = dqkv + ((long long)b * S + qi) * (3LL * NH * HD) + (long long)h * HD + g * 64;
        if (rope_cos)  // inverse RoPE of dq fused into the store (position = query index)
            store_acc_64_irope(lane_base + 256, g, dqp, scale, row_ok, rope_cos + (long long)(row_ok ? qi : 0) * HD,
                               rope_sin + (long long)(row_ok ? qi : 0) * HD);
        else
            store_acc_64(lane_base + 256 + g * 64, dqp, scale, row_ok);
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------- dK, dV
constexpr int KV_OFF_V = Q_BYTES;                 // K_j tile at 0, V_j tile at 32 KB (128 keys each)
constexpr int KV_OFF_RING = 2 * Q_BYTES;          // Q_i / dO_i tiles of 64 queries
constexpr int KV_OFF_PT = KV_OFF_RING + RING_SLOTS * KV_BYTES;
constexpr int KV_OFF_DST = KV_OFF_PT + 2 * P_BYTES;   // P^T and dS^T are double-buffered
constexpr int KV_OFF_BAR = KV_OFF_DST + 2 * P_BYTES;
constexpr int KV_OFF_ROWV = KV_OFF_BAR + 256;     // float [2][2][64]: lse*log2e and D of the streamed queries
constexpr int KV_SMEM_BYTES = KV_OFF_ROWV + 2 * 2 * 64 * 4 + 1024;

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                        const unsigned char* __restrict__ mask, const float* __restrict__ lse,
                        const float* __restrict__ Dsum, __nv_bfloat16* __restrict__ dqkv, int S, int NH, float scale,
                        float scale_log2, const __nv_bfloat16* __restrict__ rope_cos,
                        const __nv_bfloat16* __restrict__ rope_sin LLARK_ATC_SLOPES_PARAM) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;
    uint8_t* sV = smem + KV_OFF_V;
    uint8_t* sRing = smem + KV_OFF_RING;
    uint8_t* sPT = smem + KV_OFF_PT;
    uint8_t* sdST = smem + KV_OFF_DST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + KV_OFF_BAR);
    uint64_t* kv_full = bars;                      // [1]
    uint64_t* ring_full = bars + 1;                // [RING_SLOTS]
    uint64_t* ring_empty = ring_full + RING_SLOTS; // [RING_SLOTS]
    uint64_t* sdp_full = ring_empty + RING_SLOTS;  // [2]
    uint64_t* sdp_empty = sdp_full + 2;            // [2]
    uint64_t* pds_full = sdp_empty + 2;            // [2]  P^T and dS^T of tile t written to buffers t & 1
    uint64_t* pds_done = pds_full + 2;             // [2]  dV / dK MMAs of tile t retired (buffers t & 1 free)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pds_done + 2);
    float* rowv = reinterpret_cast<float*>(smem + KV_OFF_ROWV);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = blockIdx.x;  // key tile 0 has the most query tiles and starts first
    const int h = blockIdx.y, b = blockIdx.z;
    const int k0 = kt * BQ;
    const int n_q_tiles = (S + BK - 1) / BK;
    const int t0 = k0 / BK;               // first 64-query tile that can see these keys
    const int n_it = n_q_tiles - t0;      // >= 1 because k0 < S

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        mbar_init(kv_full, 1);
        for (int i = 0; i < RING_SLOTS; ++i) {
            mbar_init(&ring_full[i], 1);
            mbar_init(&ring_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sdp_full[i], 1);
            mbar_init(&sdp_empty[i], 8);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&pds_full[i], 8);
            mbar_init(&pds_done[i], 1);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;  // S^T[2] @ 0,64   dP^T[2] @ 128,192   dV @ 256   dK @ 384

    if (warp == 0) {
        if (lane == 0) {
            const int row0 = b * S;
            const int cq = h * HD, ck = NH * HD + h * HD, cv = 2 * NH * HD + h * HD;
            mbar_arrive_expect_tx(kv_full, 2 * Q_BYTES);
            for (int c = 0; c < 2; ++c)
                for (int rh = 0; rh < 2; ++rh) {
                    tma_load_2d(sK + c * (BQ * 128) + rh * (64 * 128), &tm_qkv, kv_full, ck + c * 64, row0 + k0 + rh * 64);
                    tma_load_2d(sV + c * (BQ * 128) + rh * (64 * 128), &tm_qkv, kv_full, cv + c * 64, row0 + k0 + rh * 64);
                }
            for (int t = 0; t < n_it; ++t) {
                const int qrow = row0 + (t0 + t) * BK;
#pragma unroll
                for (int which = 0; which < 2; ++which) {  // 0: Q_t, 1: dO_t
                    const int it = 2 * t + which, slot = ring_slot(it);
                    mbar_wait(&ring_empty[slot], ring_phase(it) ^ 1);
                    uint8_t* dst = sRing + slot * KV_BYTES;
                    mbar_arrive_expect_tx(&ring_full[slot], KV_BYTES);
                    const CUtensorMap* tm = which ? &tm_do : &tm_qkv;
                    tma_load_2d(dst, tm, &ring_full[slot], cq, qrow);
                    tma_load_2d(dst + BK * 128, tm, &ring_full[slot], cq + 64, qrow);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_kk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BK >> 3) << 17) |
                                      ((uint32_t)(BQ >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) |
                                      ((uint32_t)(BQ >> 4) << 24);
            auto issue_sdp = [&](int t) {
                const int sb = t & 1;
                const int itq = 2 * t, itd = 2 * t + 1;
                mbar_wait(&sdp_empty[sb], ((uint32_t)(t >> 1) & 1) ^ 1);
                mbar_wait(&ring_full[ring_slot(itq)], ring_phase(itq));
                tc_fence_after();
                mma_kk_128x64(tmem_base + sb * BK, smem_u32(sK), BQ * 128, smem_u32(sRing + ring_slot(itq) * KV_BYTES),
                              BK * 128, idesc_kk);
                mbar_wait(&ring_full[ring_slot(itd)], ring_phase(itd));
                tc_fence_after();
                mma_kk_128x64(tmem_base + 128 + sb * BK, smem_u32(sV), BQ * 128,
                              smem_u32(sRing + ring_slot(itd) * KV_BYTES), BK * 128, idesc_kk);
                umma_commit(&sdp_full[sb]);
            };
            mbar_wait(kv_full, 0);
            issue_sdp(0);
            for (int t = 0; t < n_it; ++t) {
                if (t + 1 < n_it) issue_sdp(t + 1);
                mbar_wait(&pds_full[t & 1], (uint32_t)(t >> 1) & 1);
                tc_fence_after();
                const int itq = 2 * t, itd = 2 * t + 1;
                mma_pv_128x128(tmem_base + 256, smem_u32(sPT + (t & 1) * P_BYTES),
                               smem_u32(sRing + ring_slot(itd) * KV_BYTES), idesc_pv, t != 0);
                umma_commit(&ring_empty[ring_slot(itd)]);
                mma_pv_128x128(tmem_base + 384, smem_u32(sdST + (t & 1) * P_BYTES),
                               smem_u32(sRing + ring_slot(itq) * KV_BYTES), idesc_pv, t != 0);
                umma_commit(&ring_empty[ring_slot(itq)]);
                umma_commit(&pds_done[t & 1]);
            }
        }
    } else {
        const int cw = warp - 2;
        const int g = cw >> 2;
        const int quad = warp & 3;
        const int r_in = quad * 32 + lane;   // key row inside the tile == TMEM lane
        const int key = k0 + r_in;
        const int ct = threadIdx.x - 64;     // 0..255 inside the compute group
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const bool kvalid = key < S && (!mask || mask[(long long)b * S + key] != 0);
        const float* lse_bh = lse + ((long long)b * NH + h) * S;
        const float* d_bh = Dsum + ((long long)b * NH + h) * S;
        // lse*log2e (threads 0-63) / D (threads 64-127) of the streamed queries, fetched one tile ahead so the global
        // load latency is off the per-tile critical path
        auto fetch_rowv = [&](int t) -> float {
            if (ct >= 128 || t >= n_it) return 0.f;
            const int q = (t0 + t) * BK + (ct & 63);
            if (ct < 64) return q < S ? lse_bh[q] * LOG2E : INFINITY;  // +inf => p = 0 past the sequence end
            return q < S ? d_bh[q] : 0.f;
        };
        float rowv_next = fetch_rowv(0);
        for (int t = 0; t < n_it; ++t) {
            const int sb = t & 1;
            const int qb = (t0 + t) * BK;  // first query of the streamed tile
            float* rv = rowv + sb * 128;
            if (ct < 128) rv[ct] = rowv_next;
            rowv_next = fetch_rowv(t + 1);
            compute_group_sync();  // also orders tile t-2's readers of this buffer before the writes above (see t-1)
            mbar_wait(&sdp_full[sb], (uint32_t)(t >> 1) & 1);
            tc_fence_after();
            uint32_t sv[32], dv[32];
            tmem_ld_32x32b_x32(lane_base + sb * BK + g * 32, sv);
            tmem_ld_32x32b_x32(lane_base + 128 + sb * BK + g * 32, dv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sdp_empty[sb]);

            const int qc0 = qb + g * 32;  // query of column 0 of this half row
            const bool plain = qc0 >= k0 + quad * 32 + 31;  // warp-uniform: every query of the half row sees every key
            float pt[32], ds[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const bool ok = kvalid && (plain || qc0 + c >= key);
#ifdef LLARK_ATC_ALIBI
                const float p = ok ? ex2(fmaf(__uint_as_float(sv[c]), scale_log2,
                                             fmaf(alibi_slopes[h] * LOG2E, (float)(key - (S - 1)), -rv[g * 32 + c]))) : 0.f;
#else
                const float p = ok ? ex2(fmaf(__uint_as_float(sv[c]), scale_log2, -rv[g * 32 + c])) : 0.f;
#endif
                pt[c] = p;
                ds[c] = p * (__uint_as_float(dv[c]) - rv[64 + g * 32 + c]);
            }
            if (t >= 2) {
                mbar_wait(&pds_done[t & 1], (uint32_t)((t >> 1) - 1) & 1);  // tile t-2 retired: buffers t & 1 free
                tc_fence_after();
            }
            store_half_row(sPT + (t & 1) * P_BYTES, r_in, g, pt);
            store_half_row(sdST + (t & 1) * P_BYTES, r_in, g, ds);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pds_full[t & 1]);
        }
        mbar_wait(&pds_done[(n_it - 1) & 1], (uint32_t)((n_it - 1) >> 1) & 1);  // MMAs retire in order
        tc_fence_after();
        __nv_bfloat16* dkp = dqkv + ((long long)b * S + key) * (3LL * NH * HD) + (long long)NH * HD + (long long)h * HD +
                             g * 64;
        __nv_bfloat16* dvp = dkp + (long long)NH * HD;
        store_acc_64(lane_base + 256 + g * 64, dvp, 1.f, key < S);
        if (rope_cos)  // inverse RoPE of dk fused into the store (position = key index)
            store_acc_64_irope(lane_base + 384, g, dkp, scale, key < S, rope_cos + (long long)(key < S ? key : 0) * HD,
                               rope_sin + (long long)(key < S ? key : 0) * HD);
        else
            store_acc_64(lane_base + 384 + g * 64, dkp, scale, key < S);
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace LLARK_ATC_NS (atc)

// Returns LLARK_OK when the tcgen05 kernel was launched, LLARK_EINVAL when the shape is outside its envelope
// (the caller then uses the mma.sync kernel).
int LLARK_ATC_FWD_LAUNCH(const void* qkv, const unsigned char* key_mask, void* out, float* lse, int B, int S, int n_heads,
                       float scale LLARK_ATC_SLOPES_PARAM, cudaStream_t st) {
    using namespace LLARK_ATC_NS;
    if (S > MAX_S || (long long)B * S > 0x7fffffffLL) return LLARK_EINVAL;
    CUtensorMap tm;
    const long long ld = 3LL * n_heads * HD;
    int rc = tensor_map_2d(&tm, qkv, ld, (long long)B * S, ld, 64, 64, 2);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) {
            g_last_cuda_error = (int)e;
            return LLARK_ECUDA;
        }
        attr_set = true;
    }
    dim3 grid((S + BQ - 1) / BQ, n_heads, B);
    attn_fwd_tc_kernel<<<grid, kThreads, SMEM_BYTES, st>>>(tm, key_mask, (__nv_bfloat16*)out, lse, S, n_heads,
                                                           scale * LOG2E LLARK_ATC_SLOPES_ARG);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

// Backward (dq|dk|dv into dqkv) on tcgen05; Dsum must already hold rowsum(dO o O). LLARK_EINVAL = shape outside the
// envelope (caller falls back to the mma.sync kernels).
int LLARK_ATC_BWD_LAUNCH(const void* qkv, const unsigned char* key_mask, const void* dout, const float* lse,
                       const float* Dsum, void* dqkv, int B, int S, int n_heads, float scale, const void* rope_cos,
                       const void* rope_sin LLARK_ATC_SLOPES_PARAM, cudaStream_t st) {
    using namespace LLARK_ATC_NS;
    if ((long long)B * S > 0x7fffffffLL || S > MAX_S) return LLARK_EINVAL;
    CUtensorMap tq, td;
    const long long ld = 3LL * n_heads * HD, ldo = (long long)n_heads * HD;
    int rc = tensor_map_2d(&tq, qkv, ld, (long long)B * S, ld, 64, 64, 2);
    if (rc) return rc;
    rc = tensor_map_2d(&td, dout, ldo, (long long)B * S, ldo, 64, 64, 2);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             DQ_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM_BYTES);
        if (e != cudaSuccess) {
            g_last_cuda_error = (int)e;
            return LLARK_ECUDA;
        }
        attr_set = true;
    }
    dim3 grid((S + BQ - 1) / BQ, n_heads, B);
    attn_bwd_dkdv_tc_kernel<<<grid, kBwdThreads, KV_SMEM_BYTES, st>>>(tq, td, key_mask, lse, Dsum, (__nv_bfloat16*)dqkv,
                                                                      S, n_heads, scale, scale * LOG2E,
                                                                      (const __nv_bfloat16*)rope_cos,
                                                                      (const __nv_bfloat16*)rope_sin LLARK_ATC_SLOPES_ARG);
    attn_bwd_dq_tc_kernel<<<grid, kBwdThreads, DQ_SMEM_BYTES, st>>>(tq, td, key_mask, lse, Dsum, (__nv_bfloat16*)dqkv, S,
                                                                    n_heads, scale, scale * LOG2E,
                                                                    (const __nv_bfloat16*)rope_cos,
                                                                    (const __nv_bfloat16*)rope_sin LLARK_ATC_SLOPES_ARG);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

}  // namespace llark
