// Transform-average-concatenate for the group-communication model.
//
// Reference: groupcomm_sudormrf_v2.py:343-384 (class TAC) and :405-418
// (GC_UConvBlock.forward).  For every (b, t) column, with x_g the n = Co/G
// channels of group g and H = 3n:
//     h_g = PReLU(W1 x_g + b1)                       TAC_input   (:366-367)
//     q   = PReLU(W2 mean_g(h_g) + b2)               TAC_mean    (:369-375)
//     o_g = PReLU(W3 [h_g ; q] + b3)                 TAC_output  (:376-378)
//     out = x + GlobLN_{(n,L) per (b,g)}(o)          TAC_norm    (:381-383)
// The reference materialises three permuted copies and a concat; here a CTA
// keeps the whole column block on chip: thread = (t, g), lanes along t so that
// every global access is a coalesced 128 B row segment, weights are broadcast
// from shared memory, and the mean over groups is a shared-memory reduction.
// The kernel stores o RAW plus per-(b,g) statistics; tac_apply then writes
// x + GlobLN(o) which is both the U-ConvBlock input and its residual.
#include "common.cuh"

namespace sdr {

struct TacParams {
    const float *W1, *b1, *a1, *W2, *b2, *a2, *W3, *b3, *a3;
};

template <int NPG>
__global__ void __launch_bounds__(512)
tac_kernel(const float* __restrict__ x, TacParams p, float* __restrict__ o,
           double* __restrict__ stats, int G, int L, int t_tiles) {
    constexpr int H = 3 * NPG;
    extern __shared__ __align__(16) float sm[];
    float* sW1 = sm;                      // [H][NPG]
    float* sb1 = sW1 + H * NPG;           // [H]
    float* sW2 = sb1 + H;                 // [H][H]
    float* sb2 = sW2 + H * H;             // [H]
    float* sW3 = sb2 + H;                 // [NPG][2H]
    float* sb3 = sW3 + NPG * 2 * H;       // [NPG]  (padded to 4)
    float* sMean = sb3 + ((NPG + 3) & ~3);// [H][32]
    float* sQ = sMean + H * 32;           // [H][32]
    float* sU = sQ + H * 32;              // [NPG][32]
    float* sS = sU + NPG * 32;            // [NPG][G][32]   reduction scratch

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, g = tid >> 5;
    const int b = blockIdx.x / t_tiles;
    const int t = (blockIdx.x - b * t_tiles) * 32 + lane;
    const bool valid = t < L;

    for (int i = tid; i < H * NPG; i += nthr) sW1[i] = __ldg(p.W1 + i);
    for (int i = tid; i < H; i += nthr) { sb1[i] = __ldg(p.b1 + i); sb2[i] = __ldg(p.b2 + i); }
    for (int i = tid; i < H * H; i += nthr) sW2[i] = __ldg(p.W2 + i);
    for (int i = tid; i < NPG * 2 * H; i += nthr) sW3[i] = __ldg(p.W3 + i);
    for (int i = tid; i < NPG; i += nthr) sb3[i] = __ldg(p.b3 + i);
    const float a1 = __ldg(p.a1), a2 = __ldg(p.a2), a3 = __ldg(p.a3);

    // 1. this thread's group column
    float xv[NPG];
    const size_t rowbase = ((size_t)b * G + g) * NPG;
#pragma unroll
    for (int i = 0; i < NPG; ++i) xv[i] = valid ? __ldg(x + (rowbase + i) * L + t) : 0.f;
    __syncthreads();

    // 2. h_g = PReLU(W1 x_g + b1)
    float h[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float acc = sb1[j];
#pragma unroll
        for (int i = 0; i < NPG; i += 4) {
            const float4 w = *reinterpret_cast<const float4*>(sW1 + j * NPG + i);
            acc = fmaf(w.x, xv[i], acc); acc = fmaf(w.y, xv[i + 1], acc);
            acc = fmaf(w.z, xv[i + 2], acc); acc = fmaf(w.w, xv[i + 3], acc);
        }
        h[j] = acc >= 0.f ? acc : acc * a1;
    }

    // 3. mean over groups, NPG hidden units at a time
    const float invG = 1.0f / (float)G;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int jj = 0; jj < NPG; ++jj) sS[(jj * G + g) * 32 + lane] = h[c * NPG + jj];
        __syncthreads();
        for (int jj = g; jj < NPG; jj += G) {
            float s = 0.f;
            for (int gg = 0; gg < G; ++gg) s += sS[(jj * G + gg) * 32 + lane];
            sMean[(c * NPG + jj) * 32 + lane] = s * invG;
        }
        __syncthreads();
    }

    // 4. q = PReLU(W2 mean + b2): hidden units dealt round-robin over the G warps
    for (int j = g; j < H; j += G) {
        float acc = sb2[j];
        for (int k = 0; k < H; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(sW2 + j * H + k);
            acc = fmaf(w.x, sMean[(k + 0) * 32 + lane], acc);
            acc = fmaf(w.y, sMean[(k + 1) * 32 + lane], acc);
            acc = fmaf(w.z, sMean[(k + 2) * 32 + lane], acc);
            acc = fmaf(w.w, sMean[(k + 3) * 32 + lane], acc);
        }
        sQ[j * 32 + lane] = acc >= 0.f ? acc : acc * a2;
    }
    __syncthreads();

    // 5. the q half of TAC_output is shared by all groups: u = W3[:, H:] q
    for (int i = g; i < NPG; i += G) {
        float acc = 0.f;
        for (int j = 0; j < H; j += 4) {
            const float4 w = *reinterpret_cast<const float4*>(sW3 + i * 2 * H + H + j);
            acc = fmaf(w.x, sQ[(j + 0) * 32 + lane], acc);
            acc = fmaf(w.y, sQ[(j + 1) * 32 + lane], acc);
            acc = fmaf(w.z, sQ[(j + 2) * 32 + lane], acc);
            acc = fmaf(w.w, sQ[(j + 3) * 32 + lane], acc);
        }
        sU[i * 32 + lane] = acc;
    }
    __syncthreads();

    // 6. o_g = PReLU(W3[:, :H] h_g + u + b3), raw store + per-(b,g) statistics
    float st_s = 0.f, st_q = 0.f;
#pragma unroll
    for (int i = 0; i < NPG; ++i) {
        float acc = sb3[i] + sU[i * 32 + lane];
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 w = *reinterpret_cast<const float4*>(sW3 + i * 2 * H + j);
            acc = fmaf(w.x, h[j], acc); acc = fmaf(w.y, h[j + 1], acc);
            acc = fmaf(w.z, h[j + 2], acc); acc = fmaf(w.w, h[j + 3], acc);
        }
        const float v = acc >= 0.f ? acc : acc * a3;
        if (valid) {
            o[(rowbase + i) * L + t] = v;
            st_s += v; st_q = fmaf(v, v, st_q);
        }
    }
    st_s = warp_sum(st_s);
    st_q = warp_sum(st_q);
    if (lane == 0) {
        atomicAdd(stats + 2 * ((size_t)b * G + g), (double)st_s);
        atomicAdd(stats + 2 * ((size_t)b * G + g) + 1, (double)st_q);
    }
}

template <int NPG>
static int launch_tac_n(const float* x, const TacParams& p, float* o, double* stats,
                        int B, int G, int L, cudaStream_t st) {
    constexpr int H = 3 * NPG;
    const size_t floats = (size_t)H * NPG + H + (size_t)H * H + H + (size_t)NPG * 2 * H +
                          ((NPG + 3) & ~3) + 2 * (size_t)H * 32 + (size_t)NPG * 32 +
                          (size_t)NPG * G * 32;
    const size_t smem = floats * sizeof(float);
    if (smem > 220 * 1024) return SDR_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(tac_kernel<NPG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return SDR_ERR_CUDA;
    const int t_tiles = (L + 31) / 32;
    const long long grid = (long long)t_tiles * B;
    if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
    tac_kernel<NPG><<<(unsigned)grid, 32 * G, smem, st>>>(x, p, o, stats, G, L, t_tiles);
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}

int launch_tac(const float* x, const float* const* params, float* o, double* stats,
               int B, int G, int n, int L, cudaStream_t st) {
    if (B <= 0 || G <= 0 || n <= 0 || L <= 0 || !params) return SDR_ERR_BAD_ARGUMENT;
    if (G > 16) return SDR_ERR_UNSUPPORTED;      // CTA = 32*G threads
    TacParams p{params[0], params[1], params[2], params[3], params[4],
                params[5], params[6], params[7], params[8]};
    switch (n) {
        case 4:  return launch_tac_n<4>(x, p, o, stats, B, G, L, st);
        case 8:  return launch_tac_n<8>(x, p, o, stats, B, G, L, st);
        case 16: return launch_tac_n<16>(x, p, o, stats, B, G, L, st);
        case 32: return launch_tac_n<32>(x, p, o, stats, B, G, L, st);
        default: return SDR_ERR_UNSUPPORTED;     // channels per group must be 4, 8, 16 or 32
    }
}

// out[bg, i, t] = x[bg, i, t] + GlobLN_{bg}(o)[bg, i, t]   (groupcomm_sudormrf_v2.py:381-383)
__global__ void __launch_bounds__(256)
tac_apply_kernel(const float* __restrict__ x, const float* __restrict__ o, NormIn nin,
                 float* __restrict__ out, int n, int L, int chunks_per_sample) {
    __shared__ SampleNorm s_norm;
    const int sample = blockIdx.x / chunks_per_sample;       // sample = b*G + g
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x == 0) s_norm = sample_norm(nin, sample);
    __syncthreads();
    const SampleNorm sn = s_norm;
    const int items = n * L;
    const size_t base = (size_t)sample * items;
    for (int it = 0; it < 4; ++it) {
        const int item = (chunk * 4 + it) * 256 + threadIdx.x;
        if (item < items) {
            const ChanNorm cn = chan_norm(nin, sn, item / L);
            out[base + item] = __ldg(x + base + item) + apply_norm(cn, __ldg(o + base + item));
        }
    }
}

int launch_tac_apply(const float* x, const float* o, const NormIn& nin, float* out,
                     int samples, int n, int L, cudaStream_t st) {
    const long long items = (long long)n * L;
    const int chunks = (int)((items + 1023) / 1024);
    const long long grid = (long long)chunks * samples;
    if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
    tac_apply_kernel<<<(unsigned)grid, 256, 0, st>>>(x, o, nin, out, n, L, chunks);
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}

}  // namespace sdr
