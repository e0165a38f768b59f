// fp32 verification path of the Llama forward (north_star: "fp32 logits within 1e-3 relative" of the reference's own
// PyTorch path, m2t/models/llamav2.py:259-337 -> transformers LlamaModel in fp32).
//
// Forward only, fp32 I/O everywhere. The linears run on the tcgen05 GEMM core (kind::tf32) composed as 3xTF32 on the
// host side (llark_b200/verify.py): x = hi + lo with hi = x truncated to the 10-bit TF32 mantissa (llark_v32_split_tf32),
// C = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi accumulated in fp32 through the GEMM's residual input, which recovers ~fp32
// products. Everything that is not a GEMM is one of the plain fp32 kernels below; they follow the rounding-free fp32
// formulas of the HF modules (LlamaRMSNorm, rotate_half RoPE, eager softmax attention with causal + key-padding mask,
// SiLU(gate) * up, shifted mean cross-entropy). These kernels are for checking, not for speed: small grids, no tiling.
#include "common.cuh"

namespace llark {
namespace v32 {

static int check_launch() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? LLARK_OK : LLARK_ECUDA;
}

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = l < nw ? red[l] : 0.f;
    t = warp_sum(t);
    return t;
}
__device__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = l < nw ? red[l] : -INFINITY;
    t = warp_max(t);
    return t;
}

// hi = x with the 13 low mantissa bits cleared (exactly representable in TF32, so the tensor core reads it unchanged
// whether it truncates or rounds), lo = x - hi (exact in fp32).
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                  long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float v = x[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}

// same selection rule as embed_splice_fwd_kernel (llama_ops.cu), fp32 rows
__global__ void embed_splice_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                        const float* __restrict__ audio, const long long* __restrict__ audio_off,
                                        const int* __restrict__ frames, const int* __restrict__ start_pos,
                                        float* __restrict__ out, int S, int H, long long vocab) {
    const int b = blockIdx.y, s = blockIdx.x;
    const int pos = start_pos ? start_pos[b] : -1;
    const int F = frames ? frames[b] : 0;
    const float* src;
    if (pos >= 0 && s > pos && s <= pos + F) {
        src = audio + (audio_off[b] + (long long)(s - pos - 1)) * H;
    } else {
        long long t = ids[(long long)b * S + s];
        if (t < 0 || t >= vocab) t = 0;
        src = table + t * H;
    }
    float* dst = out + ((long long)b * S + s) * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) dst[i] = src[i];
}

// LlamaRMSNorm in fp32: y = w * (x * rsqrt(mean(x^2) + eps))
__global__ void rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int H,
                                   float eps) {
    __shared__ float red[32];
    const float* xr = x + (long long)blockIdx.x * H;
    float ss = 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) ss += xr[i] * xr[i];
    ss = block_sum(ss, red);
    const float rstd = rsqrtf(ss / (float)H + eps);
    for (int i = threadIdx.x; i < H; i += blockDim.x) y[(long long)blockIdx.x * H + i] = w[i] * (xr[i] * rstd);
}

// rotate_half RoPE on the first `n_rot_heads` heads of every row of qkv [rows, ld]; position = row % S;
// cos/sin fp32 [S, D] (the HF tables: cat(freqs, freqs)). One CTA per row.
__global__ void rope_kernel(float* __restrict__ qkv, const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                            int S, int n_rot_heads, int D, long long ld) {
    const long long row = blockIdx.x;
    const int pos = (int)(row % S);
    const int half = D / 2;
    float* r = qkv + row * ld;
    for (int i = threadIdx.x; i < n_rot_heads * half; i += blockDim.x) {
        const int h = i / half, d = i % half;
        float* p = r + h * D;
        const float a = p[d], b = p[d + half];
        const float c0 = cos_t[pos * D + d], s0 = sin_t[pos * D + d];
        const float c1 = cos_t[pos * D + d + half], s1 = sin_t[pos * D + d + half];
        p[d] = a * c0 - b * s0;          // q*cos + rotate_half(q)*sin, first half: rotate_half = -x2
        p[d + half] = b * c1 + a * s1;   // second half: rotate_half = x1
    }
}

__global__ void swiglu_fwd_kernel(const float* __restrict__ gu, float* __restrict__ act, long long rows, int I) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = rows * I, stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const long long r = i / I;
        const int c = (int)(i % I);
        const float g = gu[r * 2 * I + c], u = gu[r * 2 * I + I + c];
        act[i] = (g / (1.f + expf(-g))) * u;
    }
}

// Eager attention, one warp per (b, h, query): softmax_j<=q, mask[j] (q.k_j * scale) . v in fp32. Two passes over the
// keys (max, then exp-sum and P.V); every lane owns the output columns d = lane, lane+32, ... A query whose keys are all
// masked writes zeros (the reference produces an arbitrary uniform average there; tests compare valid rows only).
__global__ void attn_fwd_kernel(const float* __restrict__ qkv, const unsigned char* __restrict__ key_mask,
                                float* __restrict__ out, int B, int S, int NH, int D, float scale) {
    extern __shared__ float sm[];  // per warp: q[D]
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * warps + w;
    if (gw >= (long long)B * NH * S) return;
    const int q_idx = (int)(gw % S);
    const int h = (int)((gw / S) % NH);
    const int b = (int)(gw / ((long long)S * NH));
    const int Hh = NH * D;
    const long long ld = 3LL * Hh;
    const float* qrow = qkv + ((long long)b * S + q_idx) * ld + h * D;
    float* qs = sm + w * D;
    for (int d = lane; d < D; d += 32) qs[d] = qrow[d];
    __syncwarp();
    const unsigned char* km = key_mask ? key_mask + (long long)b * S : nullptr;
    float mx = -INFINITY;
    for (int j = lane; j <= q_idx; j += 32) {
        if (km && !km[j]) continue;
        const float* kr = qkv + ((long long)b * S + j) * ld + Hh + h * D;
        float s = 0.f;
        for (int d = 0; d < D; ++d) s += qs[d] * kr[d];
        mx = fmaxf(mx, s * scale);
    }
    mx = warp_max(mx);
    float acc[8];  // D <= 256
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    float denom = 0.f;
    if (mx > -INFINITY) {
        for (int j0 = 0; j0 <= q_idx; j0 += 32) {
            const int j = j0 + lane;
            float p = 0.f;
            if (j <= q_idx && !(km && !km[j])) {
                const float* kr = qkv + ((long long)b * S + j) * ld + Hh + h * D;
                float s = 0.f;
                for (int d = 0; d < D; ++d) s += qs[d] * kr[d];
                p = expf(s * scale - mx);
            }
            denom += p;
            const int cnt = min(32, q_idx - j0 + 1);
            for (int t = 0; t < cnt; ++t) {
                const float pt = __shfl_sync(0xffffffffu, p, t);
                if (pt == 0.f) continue;
                const float* vr = qkv + ((long long)b * S + j0 + t) * ld + 2 * Hh + h * D;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int d = lane + 32 * i;
                    if (d < D) acc[i] += pt * vr[d];
                }
            }
        }
        denom = warp_sum(denom);
    }
    float* orow = out + ((long long)b * S + q_idx) * Hh + h * D;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int d = lane + 32 * i;
        if (d < D) orow[d] = denom > 0.f ? acc[i] / denom : 0.f;
    }
}

// Shifted cross-entropy (llamav2.py:314-325): row (b, s) predicts labels[b, s+1]; per-row loss into row_loss (0 for
// ignored rows), then a single-CTA reduction to loss_out[0] = mean over non-ignored rows, loss_out[1] = their count.
__global__ void ce_rows_kernel(const float* __restrict__ logits, long long ldv, int V, const long long* __restrict__ labels,
                               int S, float* __restrict__ row_loss) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const int s = (int)(row % S);
    long long lab = -100;
    if (s + 1 < S) lab = labels[row + 1];
    if (lab < 0 || lab >= V) {
        if (threadIdx.x == 0) row_loss[row] = -1.f;  // marker: ignored
        return;
    }
    const float* lr = logits + row * ldv;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, lr[i]);
    mx = block_max(mx, red);
    float se = 0.f;
    for (int i = threadIdx.x; i < V; i += blockDim.x) se += expf(lr[i] - mx);
    se = block_sum(se, red);
    if (threadIdx.x == 0) row_loss[row] = (logf(se) + mx) - lr[lab];
}
__global__ void ce_mean_kernel(const float* __restrict__ row_loss, long long rows, float* __restrict__ loss_out) {
    __shared__ float red[32];
    float s = 0.f, c = 0.f;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x) {
        const float v = row_loss[i];
        if (v >= 0.f) {
            s += v;
            c += 1.f;
        }
    }
    s = block_sum(s, red);
    c = block_sum(c, red);
    if (threadIdx.x == 0) {
        loss_out[0] = c > 0.f ? s / c : nanf("");
        loss_out[1] = c;
    }
}

}  // namespace v32
}  // namespace llark

using namespace llark::v32;

static int ew_blocks(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

extern "C" int llark_v32_split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t st) {
    if (n <= 0) return LLARK_EINVAL;
    split_tf32_kernel<<<ew_blocks(n), 256, 0, st>>>(x, hi, lo, n);
    return check_launch();
}
extern "C" int llark_v32_embed_splice_fwd(const long long* ids, const float* table, const float* audio,
                                          const long long* audio_off, const int* frames, const int* start_pos, float* out,
                                          int B, int S, int H, long long vocab, cudaStream_t st) {
    if (B <= 0 || S <= 0 || H <= 0) return LLARK_EINVAL;
    dim3 grid(S, B);
    embed_splice_fwd_kernel<<<grid, 128, 0, st>>>(ids, table, audio, audio_off, frames, start_pos, out, S, H, vocab);
    return check_launch();
}
extern "C" int llark_v32_rmsnorm_fwd(const float* x, const float* w, float* y, long long rows, int H, float eps,
                                     cudaStream_t st) {
    if (rows <= 0 || H <= 0) return LLARK_EINVAL;
    rmsnorm_fwd_kernel<<<(unsigned)rows, 256, 0, st>>>(x, w, y, H, eps);
    return check_launch();
}
extern "C" int llark_v32_rope(float* qkv, const float* cos_t, const float* sin_t, long long rows, int S, int n_rot_heads,
                              int D, long long ld, cudaStream_t st) {
    if (rows <= 0 || S <= 0 || D <= 0 || (D & 1)) return LLARK_EINVAL;
    rope_kernel<<<(unsigned)rows, 128, 0, st>>>(qkv, cos_t, sin_t, S, n_rot_heads, D, ld);
    return check_launch();
}
extern "C" int llark_v32_swiglu_fwd(const float* gu, float* act, long long rows, int I, cudaStream_t st) {
    if (rows <= 0 || I <= 0) return LLARK_EINVAL;
    swiglu_fwd_kernel<<<ew_blocks(rows * I), 256, 0, st>>>(gu, act, rows, I);
    return check_launch();
}
extern "C" int llark_v32_attn_fwd(const float* qkv, const unsigned char* key_mask, float* out, int B, int S, int n_heads,
                                  int D, float scale, cudaStream_t st) {
    if (B <= 0 || S <= 0 || n_heads <= 0 || D <= 0 || D > 256) return LLARK_EINVAL;
    const int warps = 4;
    const long long total = (long long)B * n_heads * S;
    attn_fwd_kernel<<<(unsigned)((total + warps - 1) / warps), warps * 32, warps * D * sizeof(float), st>>>(
        qkv, key_mask, out, B, S, n_heads, D, scale);
    return check_launch();
}
extern "C" int llark_v32_ce_fwd(const float* logits, long long ldv, int V, const long long* labels, int B, int S,
                                float* row_loss, float* loss_out, cudaStream_t st) {
    if (B <= 0 || S <= 0 || V <= 0) return LLARK_EINVAL;
    const long long rows = (long long)B * S;
    ce_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(logits, ldv, V, labels, S, row_loss);
    ce_mean_kernel<<<1, 1024, 0, st>>>(row_loss, rows, loss_out);
    return check_launch();
}
