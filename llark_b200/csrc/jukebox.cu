// Jukebox front end kernels (everything of the 5b VQ-VAE encoder + prior that is not a GEMM).
// Reference call sites: jukebox/main.py:54-68 (get_z -> vqvae.encode), :101-110 (prior.forward, only_encode),
// :113-130 (windowed_average). Arithmetic follows openai/jukebox@08efbbc as restated in SURVEY Appendix A
// (the upstream source is not vendored; oracle/jukebox_oracle.py is the CPU restatement these are tested against).
//
// All activations fp32 (the reference runs the prior with fp16=False). The conv stack is written time-tiled with
// all channels of a tile resident in shared memory, so each ResConv1DBlock is ONE pass over HBM
// (read x, write x'): the per-res-block fusion SURVEY §8(d) counts 1.42 GB/clip against.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace llark {

extern thread_local int g_last_cuda_error;
static inline int jb_check() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

// Register-tiled direct convolution from a shared-memory input tile. 256 threads = 8 warps; warp w owns the channel
// group cg = w % NCG (CPT = COUT / NCG output channels) and the time group tg = w / NCG; lane l of it computes the
// TPT output positions tg*32*TPT + l + 32*k. Weights sit in shared memory transposed to [ci][tap][co] so a thread
// fetches its CPT channels with CPT/4 broadcast LDS.128 per tap: per input channel that is TPT*KT scalar LDS +
// KT*CPT/4 LDS.128 against TPT*KT*CPT FMAs (>= 4 FMAs per shared-memory wavefront; the first version issued one
// LDS per FMA and was MIO-bound at 29 % FMA utilisation, profiles/r1h_jukebox_full.md). The accumulation order
// (bias, then ci ascending, taps ascending) is unchanged, so results are bit-identical to that version.
template <int COUT, int KT, int STRIDE, int NCG, int TPT>
struct ConvTile {
    static constexpr int CPT = COUT / NCG;
    static constexpr int TGROUPS = 8 / NCG;
    static constexpr int TILE_T = TGROUPS * 32 * TPT;  // output positions per CTA
    static_assert(CPT % 4 == 0 && 8 % NCG == 0, "tile shape");

    // sx: [c_in][in_w] input tile whose column j holds the input needed by output position 0 tap 0 at j = 0.
    // swT: [c_in][KT][COUT].
    __device__ static __forceinline__ void run(const float* __restrict__ sx, int in_w, int dil,
                                               const float* __restrict__ swT, int c_in,
                                               const float* __restrict__ bias, float (&acc)[TPT][CPT], int& cg,
                                               int& tl0) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        cg = warp % NCG;
        tl0 = (warp / NCG) * (32 * TPT) + lane;
#pragma unroll
        for (int k = 0; k < TPT; ++k)
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[k][c] = bias[cg * CPT + c];
        for (int ci = 0; ci < c_in; ++ci) {
            float xv[TPT][KT];
#pragma unroll
            for (int k = 0; k < TPT; ++k)
#pragma unroll
                for (int j = 0; j < KT; ++j) xv[k][j] = sx[ci * in_w + (tl0 + 32 * k) * STRIDE + j * dil];
#pragma unroll
            for (int j = 0; j < KT; ++j) {
                const float4* wp = reinterpret_cast<const float4*>(swT + (ci * KT + j) * COUT + cg * CPT);
#pragma unroll
                for (int c4 = 0; c4 < CPT / 4; ++c4) {
                    const float4 w = wp[c4];
#pragma unroll
                    for (int k = 0; k < TPT; ++k) {
                        acc[k][4 * c4 + 0] = fmaf(w.x, xv[k][j], acc[k][4 * c4 + 0]);
                        acc[k][4 * c4 + 1] = fmaf(w.y, xv[k][j], acc[k][4 * c4 + 1]);
                        acc[k][4 * c4 + 2] = fmaf(w.z, xv[k][j], acc[k][4 * c4 + 2]);
                        acc[k][4 * c4 + 3] = fmaf(w.w, xv[k][j], acc[k][4 * c4 + 3]);
                    }
                }
            }
        }
    }
};
// Input tile [c_in][in_w] <- x[n, :, t_first + j] (zero outside [0, T)), optionally through ReLU. Eight independent
// global loads are issued per thread before the first is consumed: the original one-load-then-store loop (with an
// integer division per element) left a single load in flight per thread and was 36 % of the res-block's stall samples.
template <bool RELU>
__device__ __forceinline__ void load_x_tile(float* sx, const float* __restrict__ xn, int c_in, int T, int in_w,
                                            int t_first) {
    for (int j = threadIdx.x; j < in_w; j += blockDim.x) {
        const int ti = t_first + j;
        const bool ok = ti >= 0 && ti < T;
        const float* src = xn + (ok ? ti : 0);
        int ci = 0;
        for (; ci + 8 <= c_in; ci += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = ok ? src[(long long)(ci + u) * T] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) sx[(ci + u) * in_w + j] = RELU ? fmaxf(v[u], 0.f) : v[u];
        }
        for (; ci < c_in; ++ci) {
            const float v = ok ? src[(long long)ci * T] : 0.f;
            sx[ci * in_w + j] = RELU ? fmaxf(v, 0.f) : v;
        }
    }
}
// global w [COUT][c_in][KT] -> shared [c_in][KT][COUT]
template <int COUT, int KT>
__device__ __forceinline__ void load_weights_T(float* swT, const float* __restrict__ w, int c_in) {
    // consecutive threads write consecutive shared-memory words (conflict-free) and gather from global/L1: the
    // transposing direction must be the read, a strided smem store is a COUT-way bank conflict
    for (int i = threadIdx.x; i < COUT * c_in * KT; i += blockDim.x) {
        const int r = i / COUT, co = i % COUT;  // r = ci*KT + k
        swT[i] = w[co * (c_in * KT) + r];
    }
}

// ---------------------------------------------------------------------------------------------- down conv k4 s2 p1
// y[n,co,t] = b[co] + sum_ci sum_k w[co,ci,k] * x[n,ci,2t-1+k].
template <int COUT>
struct DownCfg {
    using Tile = ConvTile<COUT, 4, 2, COUT / 16, 2>;
};
template <int COUT>
__global__ void __launch_bounds__(256)
conv_down_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ y, int c_in, int T) {
    using Tile = typename DownCfg<COUT>::Tile;
    extern __shared__ __align__(16) float sm[];
    const int To = T / 2;
    const int n = blockIdx.y;
    const int t0 = blockIdx.x * Tile::TILE_T;
    const int in_w = 2 * Tile::TILE_T + 2;
    float* swT = sm;                         // [c_in][4][COUT]  (first: keeps it 16-byte aligned)
    float* sx = sm + COUT * c_in * 4;        // [c_in][in_w]
    load_x_tile<false>(sx, x + (long long)n * c_in * T, c_in, T, in_w, 2 * t0 - 1);
    load_weights_T<COUT, 4>(swT, w, c_in);
    __syncthreads();
    float acc[2][Tile::CPT];
    int cg, tl0;
    Tile::run(sx, in_w, 1, swT, c_in, bias, acc, cg, tl0);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int t = t0 + tl0 + 32 * k;
        if (t < To) {
#pragma unroll
            for (int c = 0; c < Tile::CPT; ++c) y[((long long)n * COUT + cg * Tile::CPT + c) * To + t] = acc[k][c];
        }
    }
}

// ---------------------------------------------------------------------------------------------- fused res block
// y = x + W2 * relu(W1 (*dilation d, k=3) relu(x) + b1) + b2,   32 channels.
using ResTile1 = ConvTile<32, 3, 1, 2, 2>;  // 256 output positions per CTA
using ResTile2 = ConvTile<32, 1, 1, 2, 2>;
constexpr int RES_TT = ResTile1::TILE_T;
__global__ void __launch_bounds__(256)
resblock_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ y, int T, int d) {
    extern __shared__ __align__(16) float sm[];
    constexpr int C = 32;
    const int n = blockIdx.y;
    const int t0 = blockIdx.x * RES_TT;
    const int in_w = RES_TT + 2 * d;
    float* sw1T = sm;                  // [C][3][C]
    float* sw2T = sw1T + C * C * 3;    // [C][C]
    float* sh = sw2T + C * C;          // relu(h) tile [C][RES_TT]
    float* sx = sh + C * RES_TT;       // relu(x) tile [C][in_w]
    load_x_tile<true>(sx, x + (long long)n * C * T, C, T, in_w, t0 - d);
    load_weights_T<C, 3>(sw1T, w1, C);
    load_weights_T<C, 1>(sw2T, w2, C);
    __syncthreads();
    float acc[2][ResTile1::CPT];
    int cg, tl0;
    ResTile1::run(sx, in_w, d, sw1T, C, b1, acc, cg, tl0);
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int c = 0; c < ResTile1::CPT; ++c) sh[(cg * ResTile1::CPT + c) * RES_TT + tl0 + 32 * k] = fmaxf(acc[k][c], 0.f);
    __syncthreads();
    ResTile2::run(sh, RES_TT, 1, sw2T, C, b2, acc, cg, tl0);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int t = t0 + tl0 + 32 * k;
        if (t < T) {
#pragma unroll
            for (int c = 0; c < ResTile2::CPT; ++c) {
                const long long off = ((long long)n * C + cg * ResTile2::CPT + c) * T + t;
                y[off] = x[off] + acc[k][c];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- conv k3 s1 p1
template <int COUT>
struct K3Cfg {
    using Tile = ConvTile<COUT, 3, 1, COUT / 16, 2>;
};
template <int COUT>
__global__ void __launch_bounds__(256)
conv_k3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ y, int c_in, int T) {
    using Tile = typename K3Cfg<COUT>::Tile;
    extern __shared__ __align__(16) float sm[];
    const int n = blockIdx.y;
    const int t0 = blockIdx.x * Tile::TILE_T;
    const int in_w = Tile::TILE_T + 2;
    float* swT = sm;                       // [c_in][3][COUT]
    float* sx = sm + COUT * c_in * 3;      // [c_in][in_w]
    load_x_tile<false>(sx, x + (long long)n * c_in * T, c_in, T, in_w, t0 - 1);
    load_weights_T<COUT, 3>(swT, w, c_in);
    __syncthreads();
    float acc[2][Tile::CPT];
    int cg, tl0;
    Tile::run(sx, in_w, 1, swT, c_in, bias, acc, cg, tl0);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int t = t0 + tl0 + 32 * k;
        if (t < T) {
#pragma unroll
            for (int c = 0; c < Tile::CPT; ++c) y[((long long)n * COUT + cg * Tile::CPT + c) * T + t] = acc[k][c];
        }
    }
}

// ---------------------------------------------------------------------------------------------- VQ nearest code
// One thread per (n,t); x vector in registers, codebook streamed through shared memory (broadcast reads).
// dist = (|x|^2 - 2 x.k) + |k|^2, first minimum wins.
__global__ void __launch_bounds__(128)
vq_encode_kernel(const float* __restrict__ x, const float* __restrict__ cb, long long* __restrict__ codes, int T,
                 int n_codes) {
    constexpr int E = 64, CHUNK = 128;
    __shared__ float scb[CHUNK * E];
    __shared__ float skk[CHUNK];
    const int n = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float xv[E];
    float xx = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        xv[e] = (t < T) ? x[((long long)n * E + e) * T + t] : 0.f;
        xx = fmaf(xv[e], xv[e], xx);
    }
    float best = INFINITY;
    int best_j = 0;
    for (int c0 = 0; c0 < n_codes; c0 += CHUNK) {
        __syncthreads();
        for (int i = threadIdx.x; i < CHUNK * E; i += blockDim.x) {
            const int j = c0 + i / E;
            scb[i] = j < n_codes ? cb[(long long)c0 * E + i] : 0.f;
        }
        __syncthreads();
        if (threadIdx.x < CHUNK) {
            float kk = 0.f;
            for (int e = 0; e < E; ++e) kk = fmaf(scb[threadIdx.x * E + e], scb[threadIdx.x * E + e], kk);
            skk[threadIdx.x] = kk;
        }
        __syncthreads();
        const int lim = min(CHUNK, n_codes - c0);
        for (int j = 0; j < lim; ++j) {
            float dot = 0.f;
#pragma unroll
            for (int e = 0; e < E; ++e) dot = fmaf(xv[e], scb[j * E + e], dot);
            const float dist = (xx - 2.f * dot) + skk[j];
            if (dist < best) {
                best = dist;
                best_j = c0 + j;
            }
        }
    }
    if (t < T) codes[(long long)n * T + t] = best_j;
}

// ---------------------------------------------------------------------------------------------- LayerNorm
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 void* __restrict__ y, int y_bf16, int W, float eps) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const float* xr = x + row * W;
    float s = 0.f;
    for (int i = threadIdx.x * 4; i < W; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        s += v.x + v.y + v.z + v.w;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    float tot = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    tot = warp_sum(tot);
    __shared__ float s_mean, s_rstd;
    if (threadIdx.x == 0) s_mean = tot / (float)W;
    __syncthreads();
    const float mean = s_mean;
    float q = 0.f;
    for (int i = threadIdx.x * 4; i < W; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        q += a * a + b * b + c * c + d * d;
    }
    q = warp_sum(q);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    float qt = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    qt = warp_sum(qt);
    if (threadIdx.x == 0) s_rstd = rsqrtf(qt / (float)W + eps);
    __syncthreads();
    const float rstd = s_rstd;
    for (int i = threadIdx.x * 4; i < W; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        const float4 g = *reinterpret_cast<const float4*>(gamma + i);
        const float4 b = *reinterpret_cast<const float4*>(beta + i);
        const float o0 = (v.x - mean) * rstd * g.x + b.x, o1 = (v.y - mean) * rstd * g.y + b.y,
                    o2 = (v.z - mean) * rstd * g.z + b.z, o3 = (v.w - mean) * rstd * g.w + b.w;
        if (y_bf16) {
            uint2 o;
            o.x = pack_bf16x2(o0, o1);
            o.y = pack_bf16x2(o2, o3);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * W + i) = o;
        } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * W + i) = make_float4(o0, o1, o2, o3);
        }
    }
}

// ---------------------------------------------------------------------------------------------- prior input embedding
__global__ void jb_embed_kernel(const long long* __restrict__ codes, const float* __restrict__ emb,
                                const float* __restrict__ y_cond, const float* __restrict__ pos_emb,
                                const float* __restrict__ x_cond, float* __restrict__ out, int T, int W) {
    const int t = blockIdx.x, n = blockIdx.y;
    const float* src = (t == 0) ? y_cond : emb + codes[(long long)n * T + t - 1] * W;
    const float* pe = pos_emb + (long long)t * W;
    const float* xc = x_cond + (long long)t * W;
    float* dst = out + ((long long)n * T + t) * W;
    for (int i = threadIdx.x * 4; i < W; i += blockDim.x * 4) {
        const float4 a = *reinterpret_cast<const float4*>(src + i);
        const float4 b = *reinterpret_cast<const float4*>(pe + i);
        const float4 c = *reinterpret_cast<const float4*>(xc + i);
        // reference order: (x_emb + pos_emb) + x_cond
        *reinterpret_cast<float4*>(dst + i) =
            make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
    }
}

// ---------------------------------------------------------------------------------------------- factored attention
// One CTA per (group, head, n). A group is <= 128 queries and <= 128 keys gathered with a stride:
//   pattern 0 (block):            group g = block g; queries = keys = positions g*blk + i;           causal
//   pattern 1 (transposed block): group g = offset g; queries = keys = positions i*blk + g;          causal in i
//   pattern 2 (previous block):   group g = block g; queries g*blk + i, keys (g-1)*blk + i (none for g=0); full
// Each warp processes one query at a time: lanes over keys for QK^T, lanes over channels for PV (fp32 SIMT).
__global__ void __launch_bounds__(256)
jb_sparse_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int heads, int dh, int blk,
                      int pattern, float scale) {
    extern __shared__ float sm[];
    const int g = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    const int nblk = T / blk;
    int nq, nk, qstride, q0, k0;
    bool causal;
    if (pattern == 0) { nq = nk = blk; qstride = 1; q0 = k0 = g * blk; causal = true; }
    else if (pattern == 1) { nq = nk = nblk; qstride = blk; q0 = k0 = g; causal = true; }
    else { nq = blk; nk = (g == 0) ? 0 : blk; qstride = 1; q0 = g * blk; k0 = (g - 1) * blk; causal = false; }
    const int ldk = dh + 1;
    float* sK = sm;               // [nk][ldk]
    float* sV = sK + 128 * ldk;   // [nk][ldk]
    float* sQ = sV + 128 * ldk;   // [8 warps][dh]
    float* sP = sQ + 8 * dh;      // [8 warps][128]
    const long long rs = 3LL * heads * dh;
    const float* base = qkv + (long long)n * T * rs;
    for (int i = threadIdx.x; i < nk * dh; i += blockDim.x) {
        const int r = i / dh, c = i % dh;
        const long long pos = k0 + (long long)r * qstride;
        sK[r * ldk + c] = base[pos * rs + (long long)heads * dh + h * dh + c];
        sV[r * ldk + c] = base[pos * rs + 2LL * heads * dh + h * dh + c];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* myq = sQ + warp * dh;
    float* myp = sP + warp * 128;
    for (int qi = warp; qi < nq; qi += 8) {
        const long long qpos = q0 + (long long)qi * qstride;
        for (int c = lane; c < dh; c += 32) myq[c] = base[qpos * rs + h * dh + c];
        __syncwarp();
        const int kmax = causal ? qi + 1 : nk;  // keys [0, kmax)
        float sc[4];
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int kj = lane + r * 32;
            float dot = -INFINITY;
            if (kj < kmax) {
                dot = 0.f;
                const float* kr = sK + kj * ldk;
                for (int c = 0; c < dh; ++c) dot = fmaf(myq[c], kr[c], dot);
                dot *= scale;
            }
            sc[r] = dot;
            mx = fmaxf(mx, dot);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float p = (sc[r] == -INFINITY) ? 0.f : __expf(sc[r] - mx);
            sc[r] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) myp[lane + r * 32] = sc[r] * inv;
        __syncwarp();
        float* orow = out + ((long long)n * T + qpos) * ((long long)heads * dh) + h * dh;
        for (int c = lane; c < dh; c += 32) {
            float acc = 0.f;
            for (int kj = 0; kj < kmax; ++kj) acc = fmaf(myp[kj], sV[kj * ldk + c], acc);
            orow[c] = acc;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- factored attention, tensor cores
// Same contract as jb_sparse_attn_kernel, on mma.sync m16n8k8 TF32 (fp32 accumulate): the SIMT kernel above issues one
// shared-memory load per FMA and runs at ~6 TFLOP/s (4.7 ms per prior layer at batch 8, a third of the front end,
// profiles/r1_session2_summary.md); here one CTA = one (group, head, n), each warp owns 16 query rows:
//   S = Q K^T  (Q fragments in registers, K from shared memory, 19 k-steps of 8 over the 150(+2 zero) channels),
//   softmax in registers on the accumulator fragments (quad shuffles),
//   O = P V    with P taken straight from the S accumulators: the m16n8k8 A fragment wants key slots (t, t+4) where the
//              accumulator holds keys (2t, 2t+1), so the k index of the P.V product is permuted consistently -- slot t
//              reads V row 8j+2t, slot t+4 reads V row 8j+2t+1 -- and no shuffle is needed.
// K and V rows are gathered (stride 1 or `blk`) into shared memory once per CTA; while S is computed the V tile
// replaces the Q staging tile. Operands are rounded to TF32 (cvt.rna) when the fragments are formed. Causal key tiles
// above the diagonal are skipped. Output fp32 or bf16 (fused cast for the bf16 prior).
constexpr int JA_LD = 156;   // shared-memory row pitch (floats): 156 mod 32 = 28 -> conflict-free fragment loads
constexpr int JA_KS = 19;    // k-steps of 8 covering head channels (<= 152)
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// rows x dh floats gathered from global (row r at base + (pos0 + r*stride)*rs) into smem [rows][JA_LD], pad columns zero
__device__ __forceinline__ void ja_load_rows(float* dst, const float* __restrict__ base, long long pos0, int stride,
                                             long long rs, int rows, int dh) {
    constexpr int pairs = JA_LD / 2;  // float2 granules per row (dh is even, rows are 8-byte aligned)
    const int total = rows * pairs;
    // eight independent loads in flight per thread before the first shared-memory store
    for (int i0 = threadIdx.x; i0 < total; i0 += 8 * blockDim.x) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            const int r = i / pairs, c = (i % pairs) * 2;
            v[u] = make_float2(0.f, 0.f);
            if (i < total && c < dh) v[u] = *reinterpret_cast<const float2*>(base + (pos0 + (long long)r * stride) * rs + c);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            const int r = i / pairs, c = (i % pairs) * 2;
            if (i < total) *reinterpret_cast<float2*>(dst + r * JA_LD + c) = v[u];
        }
    }
}
// ---------------------------------------------------------------------------------------------- fused res block on tensor cores
// The same ResConv1DBlock  y = x + W2 . relu(W1 (*dil d, k = 3) relu(x) + b1) + b2  (32 channels) as resblock_kernel, with
// both convolutions as implicit GEMMs on mma.sync m16n8k8 TF32:  D[co, t] += A[co, k] . B[k, t],
//   conv1: k = tap * 32 + ci (K = 96, 12 k-steps),  B[k, t] = relu(x)[ci, t + (tap - 1) d]   (shifted views of ONE smem tile)
//   conv2: k = ci            (K = 32,  4 k-steps),  B[k, t] = relu(h)[ci, t]
// and fp32 accuracy kept by the 3xTF32 split: every operand is hi + lo with hi = tf32(v), lo = tf32(v - hi), and
// acc += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi (the dropped lo.lo term is ~2^-22 relative). The SIMT kernel runs at 20 TFLOP/s
// (27 % of the FP32 pipe, FMA / LDS-issue bound, profiles/r2f_prof_jukebox_hbm.md); the tensor cores take the FMAs away.
// CTA = 256 threads = 8 warps over a 256-step time tile; warp w owns time steps [32 w, 32 w + 32) (4 n-tiles of 8) and all
// 32 output channels (2 m-tiles of 16). Shared memory:
//   sx  relu(x) tile [32][LDX], LDX = in_w rounded up to 8 mod 32 so the B-fragment reads (4 rows x 8 columns per
//       instruction) touch 32 different banks;
//   sh  relu(h) / staging tile [32][LDH = 264], each warp touches only its own 32 columns (1x1 conv: no halo) -> __syncwarp;
//   weights pre-split and stored FRAGMENT-MAJOR: [k-step][m-tile][lane][a0..a3 hi, a0..a3 lo], so a thread fetches the hi and
//       lo A fragments of one (k-step, m-tile) with two conflict-free LDS.128.
constexpr int RTC_TT = 256;
constexpr int RTC_LDH = 264;
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(v);
    lo = to_tf32(v - __uint_as_float(hi));
}
// w global [32 co][32 ci][KT]; frag [(ks * 2 + mt) * 32 + lane][8]; k = tap * 32 + ci
template <int KT>
__device__ __forceinline__ void load_weight_frags(float* frag, const float* __restrict__ w) {
    constexpr int KS = KT * 4;  // k-steps of 8
    for (int i = threadIdx.x; i < KS * 2 * 32 * 4; i += blockDim.x) {
        const int r = i & 3, lane = (i >> 2) & 31, mt = (i >> 7) & 1, ks = i >> 8;
        const int g = lane >> 2, t4 = lane & 3;
        const int co = mt * 16 + g + (r & 1) * 8;
        const int k = ks * 8 + t4 + (r >> 1) * 4;
        const int tap = k >> 5, ci = k & 31;
        const float v = w[(co * 32 + ci) * KT + tap];
        uint32_t hi, lo;
        split_tf32(v, hi, lo);
        float* dst = frag + (((ks * 2 + mt) * 32 + lane) << 3);
        dst[r] = __uint_as_float(hi);
        dst[4 + r] = __uint_as_float(lo);
    }
}
__global__ void __launch_bounds__(256, 2)
resblock_tc_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                   const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ y, int T, int d,
                   int ldx) {
    extern __shared__ __align__(16) float sm[];
    constexpr int C = 32;
    const int n = blockIdx.y;
    const int t0 = blockIdx.x * RTC_TT;
    const int in_w = RTC_TT + 2 * d;
    float* f1 = sm;                       // conv1 weight fragments: 12 * 2 * 32 * 8
    float* f2 = f1 + 12 * 2 * 32 * 8;     // conv2 weight fragments:  4 * 2 * 32 * 8
    float* sh = f2 + 4 * 2 * 32 * 8;      // [C][RTC_LDH]
    float* sx = sh + C * RTC_LDH;         // [C][ldx]
    const float* xn = x + (long long)n * C * T;
    // relu(x) tile, columns t0 - d .. t0 + 255 + d (zero outside [0, T)); eight independent loads per thread
    for (int j = threadIdx.x; j < in_w; j += blockDim.x) {
        const int ti = t0 - d + j;
        const bool ok = ti >= 0 && ti < T;
        const float* src = xn + (ok ? ti : 0);
        for (int ci = 0; ci < C; ci += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = ok ? src[(long long)(ci + u) * T] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) sx[(ci + u) * ldx + j] = fmaxf(v[u], 0.f);
        }
    }
    load_weight_frags<3>(f1, w1);
    load_weight_frags<1>(f2, w2);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int tw = warp * 32;  // this warp's first time step inside the tile
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const float blo = b1[mt * 16 + g], bhi = b1[mt * 16 + g + 8];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            acc[mt][nt][0] = acc[mt][nt][1] = blo;
            acc[mt][nt][2] = acc[mt][nt][3] = bhi;
        }
    }
    // ---- conv1: 12 k-steps (tap-major), B from shifted columns of sx
#pragma unroll 1
    for (int ks = 0; ks < 12; ++ks) {
        const int tap = ks >> 2, ci0 = (ks & 3) * 8;
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float4* fp = reinterpret_cast<const float4*>(f1 + (((ks * 2 + mt) * 32 + lane) << 3));
            const float4 h = fp[0], l = fp[1];
            ah[mt][0] = __float_as_uint(h.x); ah[mt][1] = __float_as_uint(h.y);
            ah[mt][2] = __float_as_uint(h.z); ah[mt][3] = __float_as_uint(h.w);
            al[mt][0] = __float_as_uint(l.x); al[mt][1] = __float_as_uint(l.y);
            al[mt][2] = __float_as_uint(l.z); al[mt][3] = __float_as_uint(l.w);
        }
        const float* r0 = sx + (ci0 + t4) * ldx + tw + g + tap * d;
        const float* r1 = r0 + 4 * ldx;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(r0[nt * 8], bh0, bl0);
            split_tf32(r1[nt * 8], bh1, bl1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma_tf32(acc[mt][nt], al[mt], bh0, bh1);
                mma_tf32(acc[mt][nt], ah[mt], bl0, bl1);
                mma_tf32(acc[mt][nt], ah[mt], bh0, bh1);
            }
        }
    }
    // ---- relu(h) -> this warp's 32 columns of sh (accumulator layout: rows g / g + 8, columns 2 t4, 2 t4 + 1)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float* p0 = sh + (mt * 16 + g) * RTC_LDH + tw + nt * 8 + 2 * t4;
            *reinterpret_cast<float2*>(p0) = make_float2(fmaxf(acc[mt][nt][0], 0.f), fmaxf(acc[mt][nt][1], 0.f));
            *reinterpret_cast<float2*>(p0 + 8 * RTC_LDH) = make_float2(fmaxf(acc[mt][nt][2], 0.f), fmaxf(acc[mt][nt][3], 0.f));
        }
    __syncwarp();
    // ---- conv2 (1x1): 4 k-steps, B from sh
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const float blo = b2[mt * 16 + g], bhi = b2[mt * 16 + g + 8];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            acc[mt][nt][0] = acc[mt][nt][1] = blo;
            acc[mt][nt][2] = acc[mt][nt][3] = bhi;
        }
    }
#pragma unroll 1
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float4* fp = reinterpret_cast<const float4*>(f2 + (((ks * 2 + mt) * 32 + lane) << 3));
            const float4 h = fp[0], l = fp[1];
            ah[mt][0] = __float_as_uint(h.x); ah[mt][1] = __float_as_uint(h.y);
            ah[mt][2] = __float_as_uint(h.z); ah[mt][3] = __float_as_uint(h.w);
            al[mt][0] = __float_as_uint(l.x); al[mt][1] = __float_as_uint(l.y);
            al[mt][2] = __float_as_uint(l.z); al[mt][3] = __float_as_uint(l.w);
        }
        const float* r0 = sh + (ks * 8 + t4) * RTC_LDH + tw + g;
        const float* r1 = r0 + 4 * RTC_LDH;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(r0[nt * 8], bh0, bl0);
            split_tf32(r1[nt * 8], bh1, bl1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma_tf32(acc[mt][nt], al[mt], bh0, bh1);
                mma_tf32(acc[mt][nt], ah[mt], bl0, bl1);
                mma_tf32(acc[mt][nt], ah[mt], bh0, bh1);
            }
        }
    }
    __syncwarp();  // every lane has finished reading relu(h) before the tile is reused as the output staging area
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float* p0 = sh + (mt * 16 + g) * RTC_LDH + tw + nt * 8 + 2 * t4;
            *reinterpret_cast<float2*>(p0) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
            *reinterpret_cast<float2*>(p0 + 8 * RTC_LDH) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
    __syncwarp();
    // ---- y = x + (.): one channel row of this warp's 32 time steps per instruction (128 contiguous bytes)
    const int t = t0 + tw + lane;
    if (t < T) {
#pragma unroll 8
        for (int co = 0; co < C; ++co) {
            const long long off = ((long long)n * C + co) * T + t;
            y[off] = x[off] + sh[co * RTC_LDH + tw + lane];
        }
    }
}

template <int NKT>  // key tiles of 8 (8 -> 64 keys, 16 -> 128 keys)
__global__ void __launch_bounds__(256)
jb_sparse_attn_tc_kernel(const float* __restrict__ qkv, void* __restrict__ out, int out_bf16, int T, int heads, int dh,
                         int blk, int pattern, float scale) {
    extern __shared__ __align__(16) float sm[];
    constexpr int NK = NKT * 8;
    float* sK = sm;                 // [NK][JA_LD]
    float* sV = sm + NK * JA_LD;    // [NK][JA_LD]  (first holds the Q tile)
    const int g = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    int qstride, q0, k0;
    bool causal, has_keys = true;
    if (pattern == 0) { qstride = 1; q0 = k0 = g * blk; causal = true; }
    else if (pattern == 1) { qstride = blk; q0 = k0 = g; causal = true; }
    else { qstride = 1; q0 = g * blk; k0 = (g - 1) * blk; causal = false; has_keys = g > 0; }
    const long long rs = 3LL * heads * dh;
    const long long ow = (long long)heads * dh;
    const float* base = qkv + (long long)n * T * rs + (long long)h * dh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gq = lane >> 2, t4 = lane & 3;
    const int r0 = warp * 16 + gq;  // this thread's rows inside the group: r0 and r0 + 8
    if (!has_keys) {  // previous-block pattern, first block: no keys -> zeros (the SIMT kernel and the reference agree)
        for (int i = threadIdx.x; i < NK * dh; i += blockDim.x) {
            const long long pos = q0 + (long long)(i / dh) * qstride;
            const long long o = ((long long)n * T + pos) * ow + (long long)h * dh + i % dh;
            if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(0.f);
            else reinterpret_cast<float*>(out)[o] = 0.f;
        }
        return;
    }
    ja_load_rows(sK, base + (long long)heads * dh, k0, qstride, rs, NK, dh);
    ja_load_rows(sV, base, q0, qstride, rs, NK, dh);  // Q staged in the V buffer
    __syncthreads();
    uint32_t qa[JA_KS][4];
#pragma unroll
    for (int ks = 0; ks < JA_KS; ++ks) {
        qa[ks][0] = to_tf32(sV[r0 * JA_LD + ks * 8 + t4]);
        qa[ks][1] = to_tf32(sV[(r0 + 8) * JA_LD + ks * 8 + t4]);
        qa[ks][2] = to_tf32(sV[r0 * JA_LD + ks * 8 + t4 + 4]);
        qa[ks][3] = to_tf32(sV[(r0 + 8) * JA_LD + ks * 8 + t4 + 4]);
    }
    __syncthreads();  // every warp holds its Q fragments: the buffer may now receive V
    ja_load_rows(sV, base + 2LL * heads * dh, k0, qstride, rs, NK, dh);
    // key tiles this warp needs: causal rows 16w .. 16w+15 see keys <= 16w+15
    const int nt_end = causal ? min(NKT, 2 * warp + 2) : NKT;
    float sc[NKT][4];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
        if (nt < nt_end) {
            const float* kr = sK + (nt * 8 + gq) * JA_LD + t4;
#pragma unroll
            for (int ks = 0; ks < JA_KS; ++ks) mma_tf32(sc[nt], qa[ks], to_tf32(kr[ks * 8]), to_tf32(kr[ks * 8 + 4]));
        }
    }
    // softmax over the keys of rows r0 (elements 0,1) and r0+8 (elements 2,3)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int key = nt * 8 + 2 * t4 + (e & 1);
            const int row = r0 + (e >> 1) * 8;
            const bool ok = nt < nt_end && (!causal || key <= row);
            const float v = ok ? sc[nt][e] * scale : -INFINITY;
            sc[nt][e] = v;
            mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float p = (sc[nt][e] == -INFINITY) ? 0.f : __expf(sc[nt][e] - mx[e >> 1]);
            sc[nt][e] = p;
            sum[e >> 1] += p;
        }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
    }
    __syncthreads();  // V tile complete in shared memory
    float oacc[JA_KS][4];
#pragma unroll
    for (int nn = 0; nn < JA_KS; ++nn) oacc[nn][0] = oacc[nn][1] = oacc[nn][2] = oacc[nn][3] = 0.f;
#pragma unroll
    for (int j = 0; j < NKT; ++j) {
        if (j < nt_end) {
            uint32_t pa[4];
            pa[0] = to_tf32(sc[j][0]);  // (row r0,   key 8j+2t)   -> k slot t
            pa[1] = to_tf32(sc[j][2]);  // (row r0+8, key 8j+2t)
            pa[2] = to_tf32(sc[j][1]);  // (row r0,   key 8j+2t+1) -> k slot t+4
            pa[3] = to_tf32(sc[j][3]);
            const float* v0 = sV + (j * 8 + 2 * t4) * JA_LD + gq;
#pragma unroll
            for (int nn = 0; nn < JA_KS; ++nn) mma_tf32(oacc[nn], pa, to_tf32(v0[nn * 8]), to_tf32(v0[JA_LD + nn * 8]));
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const float inv = sum[r] > 0.f ? 1.f / sum[r] : 0.f;
        const long long pos = q0 + (long long)(r0 + r * 8) * qstride;
        const long long o = ((long long)n * T + pos) * ow + (long long)h * dh;
#pragma unroll
        for (int nn = 0; nn < JA_KS; ++nn) {
            const int col = nn * 8 + 2 * t4;
            if (col < dh) {
                const float a = oacc[nn][2 * r] * inv, b = oacc[nn][2 * r + 1] * inv;
                if (out_bf16)
                    *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + o + col) = pack_bf16x2(a, b);
                else
                    *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + o + col) = make_float2(a, b);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- elementwise
__global__ void quick_gelu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (; i + 4 <= n; i += stride) {
        float4 v = *reinterpret_cast<const float4*>(x + i);
        v.x = v.x / (1.f + __expf(-1.702f * v.x));
        v.y = v.y / (1.f + __expf(-1.702f * v.y));
        v.z = v.z / (1.f + __expf(-1.702f * v.z));
        v.w = v.w / (1.f + __expf(-1.702f * v.w));
        *reinterpret_cast<float4*>(y + i) = v;
    }
    if (i < n)
        for (long long j = i; j < n; ++j) y[j] = x[j] / (1.f + __expf(-1.702f * x[j]));
}
__global__ void add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                            float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = c ? (a[i] + b[i]) + c[i] : a[i] + b[i];
}
// out[n, f, c] = mean_{j<frame} acts[n, f*frame + j, c]; grid (frames, N, ceil(W/4/blockDim)), 4 columns per thread,
// the `frame` row loads of a thread are independent (two accumulators) so a warp keeps ~frame x 512 B in flight.
__global__ void jb_pool_kernel(const float* __restrict__ acts, float* __restrict__ out, long long clip_stride, int frames,
                               int W, int frame) {
    const int f = blockIdx.x;
    const int c = (blockIdx.z * blockDim.x + threadIdx.x) * 4;
    if (c >= W) return;
    const float* src = acts + (long long)blockIdx.y * clip_stride + (long long)f * frame * W + c;
    float* dst = out + ((long long)blockIdx.y * frames + f) * W + c;
    if (c + 4 <= W && (W & 3) == 0) {
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        int j = 0;
        for (; j + 1 < frame; j += 2) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + (long long)j * W));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + (long long)(j + 1) * W));
            s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
            s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
        }
        if (j < frame) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + (long long)j * W));
            s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
        }
        const float inv = 1.f / (float)frame;
        *reinterpret_cast<float4*>(dst) =
            make_float4((s0.x + s1.x) * inv, (s0.y + s1.y) * inv, (s0.z + s1.z) * inv, (s0.w + s1.w) * inv);
    } else {
        for (int k = 0; k < 4 && c + k < W; ++k) {
            float sum = 0.f;
            for (int j = 0; j < frame; ++j) sum += src[(long long)j * W + k];
            dst[k] = sum / (float)frame;
        }
    }
}
// Timing conditioning of the top-level prior (LabelConditioner.forward -> pos_emb, jukebox/main.py:93):
// out[t, :] = (total_row + abs_table[abs_bins[t], :]) + rel_table[rel_bins[t], :] -- the reference's order of additions.
__global__ void jb_label_cond_kernel(const float* __restrict__ total_row, const float* __restrict__ abs_table,
                                     const float* __restrict__ rel_table, const int* __restrict__ abs_bins,
                                     const int* __restrict__ rel_bins, float* __restrict__ out, int W) {
    const int t = blockIdx.x;
    const float* a = abs_table + (long long)abs_bins[t] * W;
    const float* r = rel_table + (long long)rel_bins[t] * W;
    float* o = out + (long long)t * W;
    for (int c = threadIdx.x; c < W; c += blockDim.x) o[c] = (total_row[c] + a[c]) + r[c];
}
// colsum over rows of a bf16 [rows, H] matrix -> bf16 [H] (bias gradient), fp32 accumulate.
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows,
                                   int H, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H) return;
    float s = 0.f;
    for (long long r = 0; r < rows; ++r) s += __bfloat162float(x[r * H + c]);
    if (accumulate) s += __bfloat162float(out[c]);
    out[c] = __float2bfloat16_rn(s);
}

static int set_dyn_smem(const void* fn, int bytes) {
    if (bytes <= 48 * 1024) return LLARK_OK;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

}  // namespace llark

using namespace llark;

extern "C" int llark_jb_conv_down(const float* x, const float* w, const float* bias, float* y, int N, int c_in,
                                  int c_out, int T, cudaStream_t st) {
    if (N <= 0 || T <= 0 || (T & 1) || c_in <= 0 || (c_out != 32 && c_out != 64)) return LLARK_EINVAL;
    int rc;
    if (c_out == 32) {
        constexpr int TILE = DownCfg<32>::Tile::TILE_T;
        const int smem = (c_in * (2 * TILE + 2) + c_out * c_in * 4) * (int)sizeof(float);
        dim3 grid((T / 2 + TILE - 1) / TILE, N);
        rc = set_dyn_smem((const void*)conv_down_kernel<32>, smem);
        if (rc) return rc;
        conv_down_kernel<32><<<grid, 256, smem, st>>>(x, w, bias, y, c_in, T);
    } else {
        constexpr int TILE = DownCfg<64>::Tile::TILE_T;
        const int smem = (c_in * (2 * TILE + 2) + c_out * c_in * 4) * (int)sizeof(float);
        dim3 grid((T / 2 + TILE - 1) / TILE, N);
        rc = set_dyn_smem((const void*)conv_down_kernel<64>, smem);
        if (rc) return rc;
        conv_down_kernel<64><<<grid, 256, smem, st>>>(x, w, bias, y, c_in, T);
    }
    return jb_check();
}

extern "C" int llark_jb_resblock(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                 float* y, int N, int T, int dilation, cudaStream_t st) {
    if (N <= 0 || T <= 0 || dilation < 1 || dilation > 243) return LLARK_EINVAL;
    const char* tc_env = getenv("LLARK_JB_CONV_TC");  // "0" selects the fp32 SIMT kernel (read per call: tests toggle it)
    const int use_tc = (tc_env && tc_env[0] == '0') ? 0 : 1;
    if (use_tc && dilation <= 81) {
        const int in_w = RTC_TT + 2 * dilation;
        const int ldx = ((in_w + 23) / 32) * 32 + 8;  // >= in_w and == 8 (mod 32): conflict-free B-fragment reads
        const int smem_tc = ((12 + 4) * 2 * 32 * 8 + 32 * RTC_LDH + 32 * ldx) * (int)sizeof(float);
        int rc_tc = set_dyn_smem((const void*)resblock_tc_kernel, smem_tc);
        if (rc_tc) return rc_tc;
        dim3 grid_tc((T + RTC_TT - 1) / RTC_TT, N);
        resblock_tc_kernel<<<grid_tc, 256, smem_tc, st>>>(x, w1, b1, w2, b2, y, T, dilation, ldx);
        return jb_check();
    }
    const int smem = (32 * (RES_TT + 2 * dilation) + 32 * RES_TT + 32 * 32 * 3 + 32 * 32) * (int)sizeof(float);
    int rc = set_dyn_smem((const void*)resblock_kernel, smem);
    if (rc) return rc;
    dim3 grid((T + RES_TT - 1) / RES_TT, N);
    resblock_kernel<<<grid, 256, smem, st>>>(x, w1, b1, w2, b2, y, T, dilation);
    return jb_check();
}

extern "C" int llark_jb_conv_k3(const float* x, const float* w, const float* bias, float* y, int N, int c_in, int c_out,
                                int T, cudaStream_t st) {
    if (N <= 0 || T <= 0 || c_in <= 0 || (c_out != 32 && c_out != 64)) return LLARK_EINVAL;
    int rc;
    if (c_out == 64) {
        constexpr int TILE = K3Cfg<64>::Tile::TILE_T;
        const int smem = (c_in * (TILE + 2) + c_out * c_in * 3) * (int)sizeof(float);
        dim3 grid((T + TILE - 1) / TILE, N);
        rc = set_dyn_smem((const void*)conv_k3_kernel<64>, smem);
        if (rc) return rc;
        conv_k3_kernel<64><<<grid, 256, smem, st>>>(x, w, bias, y, c_in, T);
    } else {
        constexpr int TILE = K3Cfg<32>::Tile::TILE_T;
        const int smem = (c_in * (TILE + 2) + c_out * c_in * 3) * (int)sizeof(float);
        dim3 grid((T + TILE - 1) / TILE, N);
        rc = set_dyn_smem((const void*)conv_k3_kernel<32>, smem);
        if (rc) return rc;
        conv_k3_kernel<32><<<grid, 256, smem, st>>>(x, w, bias, y, c_in, T);
    }
    return jb_check();
}

extern "C" int llark_jb_vq_encode(const float* x, const float* codebook, long long* codes, int N, int T, int n_codes,
                                  cudaStream_t st) {
    if (N <= 0 || T <= 0 || n_codes <= 0) return LLARK_EINVAL;
    dim3 grid((T + 127) / 128, N);
    vq_encode_kernel<<<grid, 128, 0, st>>>(x, codebook, codes, T, n_codes);
    return jb_check();
}

extern "C" int llark_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_bf16,
                                   long long rows, int W, float eps, cudaStream_t st) {
    if (rows <= 0 || W <= 0 || (W & 3)) return LLARK_EINVAL;
    layernorm_kernel<<<(unsigned)rows, 256, 0, st>>>(x, gamma, beta, y, y_bf16, W, eps);
    return jb_check();
}

extern "C" int llark_jb_embed_tokens(const long long* codes, const float* emb, const float* y_cond,
                                     const float* pos_emb, const float* x_cond, float* out, int N, int T, int W,
                                     cudaStream_t st) {
    if (N <= 0 || T <= 0 || (W & 3)) return LLARK_EINVAL;
    dim3 grid(T, N);
    jb_embed_kernel<<<grid, 256, 0, st>>>(codes, emb, y_cond, pos_emb, x_cond, out, T, W);
    return jb_check();
}

extern "C" int llark_jb_sparse_attn(const float* qkv, float* out, int N, int T, int heads, int dh, int block,
                                    int pattern, cudaStream_t st) {
    if (N <= 0 || T <= 0 || block <= 0 || T % block || pattern < 0 || pattern > 2) return LLARK_EINVAL;
    const int nblk = T / block;
    if (block > 128 || nblk > 128 || dh > 256) return LLARK_EINVAL;
    const int groups = (pattern == 1) ? block : nblk;
    const int smem = (2 * 128 * (dh + 1) + 8 * dh + 8 * 128) * (int)sizeof(float);
    int rc = set_dyn_smem((const void*)jb_sparse_attn_kernel, smem);
    if (rc) return rc;
    dim3 grid(groups, heads, N);
    jb_sparse_attn_kernel<<<grid, 256, smem, st>>>(qkv, out, T, heads, dh, block, pattern, 1.f / sqrtf((float)dh));
    return jb_check();
}

// Tensor-core variant (TF32 operands, fp32 accumulate); out fp32 or bf16. LLARK_EINVAL when the geometry is outside
// its envelope (64- or 128-row groups, even head width <= 152): callers then use llark_jb_sparse_attn.
extern "C" int llark_jb_sparse_attn_tc(const float* qkv, void* out, int out_bf16, int N, int T, int heads, int dh,
                                       int block, int pattern, cudaStream_t st) {
    if (N <= 0 || T <= 0 || block <= 0 || T % block || pattern < 0 || pattern > 2) return LLARK_EINVAL;
    const int nblk = T / block;
    const int rows = (pattern == 1) ? nblk : block;  // queries (= keys) per group
    if ((rows != 64 && rows != 128) || dh > 8 * JA_KS || (dh & 1)) return LLARK_EINVAL;
    const int groups = (pattern == 1) ? block : nblk;
    const int smem = 2 * rows * JA_LD * (int)sizeof(float);
    dim3 grid(groups, heads, N);
    const float scale = 1.f / sqrtf((float)dh);
    int rc;
    if (rows == 64) {
        rc = set_dyn_smem((const void*)jb_sparse_attn_tc_kernel<8>, smem);
        if (rc) return rc;
        jb_sparse_attn_tc_kernel<8><<<grid, 128, smem, st>>>(qkv, out, out_bf16, T, heads, dh, block, pattern, scale);
    } else {
        rc = set_dyn_smem((const void*)jb_sparse_attn_tc_kernel<16>, smem);
        if (rc) return rc;
        jb_sparse_attn_tc_kernel<16><<<grid, 256, smem, st>>>(qkv, out, out_bf16, T, heads, dh, block, pattern, scale);
    }
    return jb_check();
}

static inline int ew_grid(long long n, int per_thread) {
    long long b = (n / per_thread + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return (int)b;
}
extern "C" int llark_quick_gelu(const float* x, float* y, long long n, cudaStream_t st) {
    if (n <= 0) return LLARK_OK;
    quick_gelu_kernel<<<ew_grid(n, 4), 256, 0, st>>>(x, y, n);
    return jb_check();
}
extern "C" int llark_add3(const float* a, const float* b, const float* c, float* out, long long n, cudaStream_t st) {
    if (n <= 0) return LLARK_OK;
    add3_kernel<<<ew_grid(n, 1), 256, 0, st>>>(a, b, c, out, n);
    return jb_check();
}
extern "C" int llark_jb_pool_batched(const float* acts, float* out, int N, long long clip_stride, int T_valid, int W,
                                     int frame, cudaStream_t st) {
    if (frame <= 0 || W <= 0 || N <= 0) return LLARK_EINVAL;
    const int frames = T_valid / frame;
    if (frames <= 0) return LLARK_OK;
    dim3 grid(frames, N, (W + 1023) / 1024);
    jb_pool_kernel<<<grid, 256, 0, st>>>(acts, out, clip_stride, frames, W, frame);
    return jb_check();
}
extern "C" int llark_jb_pool(const float* acts, float* out, int T_valid, int W, int frame, cudaStream_t st) {
    return llark_jb_pool_batched(acts, out, 1, 0, T_valid, W, frame, st);
}
extern "C" int llark_jb_label_cond(const float* total_row, const float* abs_table, const float* rel_table,
                                   const int* abs_bins, const int* rel_bins, float* out, int T, int W, cudaStream_t st) {
    if (T <= 0 || W <= 0) return LLARK_EINVAL;
    jb_label_cond_kernel<<<T, 256, 0, st>>>(total_row, abs_table, rel_table, abs_bins, rel_bins, out, W);
    return jb_check();
}
extern "C" int llark_colsum(const void* x, void* out, long long rows, int H, int accumulate, cudaStream_t st) {
    if (rows <= 0 || H <= 0) return LLARK_EINVAL;
    colsum_bf16_kernel<<<(H + 127) / 128, 128, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows, H, accumulate);
    return jb_check();
}
