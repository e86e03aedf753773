// U-ConvBlock level kernels (HBM-bound): depthwise k=5 stride 1|2 with the
// producer's GlobLN(+PReLU) applied on load, and the multi-resolution merge.
//
// Reference semantics:
//   DilatedConvNorm  improved_sudormrf.py:138-159  (Conv1d k=5, pad 2, groups=C, then GlobLN)
//   UConvBlock.forward level loops  improved_sudormrf.py:206-216
// Each kernel stores its RAW result once and accumulates the (sum, sumsq) of
// that result per sample in fp64 so the consumer can normalise while loading.
#include "common.cuh"

namespace sdr {

constexpr int kDwThreads = 256;
constexpr int kDwItems = 2;        // output quads per thread

// ---------------------------------------------------------------------------
// depthwise, vector path: one item = 4 consecutive outputs of one (sample, c) row
// requires Lout % 4 == 0 (then Lin % 4 == 0 as well).
// ---------------------------------------------------------------------------
template <int STRIDE>
__global__ void __launch_bounds__(kDwThreads)
dw5_vec_kernel(const float* __restrict__ x, NormIn nin,
               const float* __restrict__ w5, const float* __restrict__ bias,
               float* __restrict__ y, double* __restrict__ stats_out,
               int C, int Lin, int Lout, int chunks_per_sample) {
    __shared__ SampleNorm s_norm;
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x == 0) s_norm = sample_norm(nin, sample);
    __syncthreads();
    const SampleNorm sn = s_norm;

    const int QR = Lout >> 2;                 // quads per row
    const int items = C * QR;                 // per sample
    const float* xs = x + (size_t)sample * C * Lin;
    float* ys = y + (size_t)sample * C * Lout;

    float acc_s = 0.f, acc_q = 0.f;
#pragma unroll
    for (int it = 0; it < kDwItems; ++it) {
        const int item = (chunk * kDwItems + it) * kDwThreads + threadIdx.x;
        if (item < items) {
            const int c = item / QR;
            const int q = item - c * QR;
            const ChanNorm cn = chan_norm(nin, sn, c);
            const float* xr = xs + (size_t)c * Lin;
            float w[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) w[j] = __ldg(w5 + c * 5 + j);
            const float b = __ldg(bias + c);
            float o[4];
            if (STRIDE == 1) {
                // window v[0..7] = positions 4q-2 .. 4q+5
                float v[8];
                const float4 m = ldg4(xr + 4 * q);
                v[2] = apply_norm(cn, m.x); v[3] = apply_norm(cn, m.y);
                v[4] = apply_norm(cn, m.z); v[5] = apply_norm(cn, m.w);
                if (q > 0) {
                    const float2 l = __ldg(reinterpret_cast<const float2*>(xr + 4 * q - 2));
                    v[0] = apply_norm(cn, l.x); v[1] = apply_norm(cn, l.y);
                } else { v[0] = 0.f; v[1] = 0.f; }
                if (q < QR - 1) {
                    const float2 r = __ldg(reinterpret_cast<const float2*>(xr + 4 * q + 4));
                    v[6] = apply_norm(cn, r.x); v[7] = apply_norm(cn, r.y);
                } else { v[6] = 0.f; v[7] = 0.f; }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float a = b;
#pragma unroll
                    for (int j = 0; j < 5; ++j) a = fmaf(w[j], v[i + j], a);
                    o[i] = a;
                }
            } else {
                // outputs 4q..4q+3 read inputs 8q-2 .. 8q+8 ; window v[0..10]
                float v[11];
                const float4 m0 = ldg4(xr + 8 * q);
                const float4 m1 = ldg4(xr + 8 * q + 4);
                v[2] = apply_norm(cn, m0.x); v[3] = apply_norm(cn, m0.y);
                v[4] = apply_norm(cn, m0.z); v[5] = apply_norm(cn, m0.w);
                v[6] = apply_norm(cn, m1.x); v[7] = apply_norm(cn, m1.y);
                v[8] = apply_norm(cn, m1.z); v[9] = apply_norm(cn, m1.w);
                if (q > 0) {
                    const float2 l = __ldg(reinterpret_cast<const float2*>(xr + 8 * q - 2));
                    v[0] = apply_norm(cn, l.x); v[1] = apply_norm(cn, l.y);
                } else { v[0] = 0.f; v[1] = 0.f; }
                v[10] = (8 * q + 8 < Lin) ? apply_norm(cn, __ldg(xr + 8 * q + 8)) : 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float a = b;
#pragma unroll
                    for (int j = 0; j < 5; ++j) a = fmaf(w[j], v[2 * i + j], a);
                    o[i] = a;
                }
            }
            *reinterpret_cast<float4*>(ys + (size_t)c * Lout + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc_s += o[i]; acc_q = fmaf(o[i], o[i], acc_q); }
        }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}

// scalar fallback (any Lin/Lout): one output element per thread-iteration
__global__ void __launch_bounds__(kDwThreads)
dw5_scalar_kernel(const float* __restrict__ x, NormIn nin,
                  const float* __restrict__ w5, const float* __restrict__ bias,
                  float* __restrict__ y, double* __restrict__ stats_out,
                  int C, int Lin, int Lout, int stride, int chunks_per_sample) {
    __shared__ SampleNorm s_norm;
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x == 0) s_norm = sample_norm(nin, sample);
    __syncthreads();
    const SampleNorm sn = s_norm;
    const int items = C * Lout;
    const float* xs = x + (size_t)sample * C * Lin;
    float* ys = y + (size_t)sample * C * Lout;
    float acc_s = 0.f, acc_q = 0.f;
    for (int it = 0; it < 4; ++it) {
        const int item = (chunk * 4 + it) * kDwThreads + threadIdx.x;
        if (item < items) {
            const int c = item / Lout;
            const int t = item - c * Lout;
            const ChanNorm cn = chan_norm(nin, sn, c);
            float a = __ldg(bias + c);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int p = t * stride + j - 2;
                if (p >= 0 && p < Lin)
                    a = fmaf(__ldg(w5 + c * 5 + j), apply_norm(cn, __ldg(xs + (size_t)c * Lin + p)), a);
            }
            ys[(size_t)c * Lout + t] = a;
            acc_s += a; acc_q = fmaf(a, a, acc_q);
        }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}

// ---------------------------------------------------------------------------
// merge: m[c,t] = sum_d norm_d(z_d)[c, t >> d]   (closed form of the
// upsample(scale 2, nearest)+add chain, improved_sudormrf.py:214-216)
// ---------------------------------------------------------------------------
constexpr int kMaxDepth = kMaxDepthApi;
struct MergeArgs {
    const float* z[kMaxDepth];
    NormIn n[kMaxDepth];
    int depth;
};

constexpr int kMgThreads = 256;
constexpr int kMgItems = 2;

// vector path: L % 4 == 0 and L % 2^(depth-1) == 0
__global__ void __launch_bounds__(kMgThreads)
merge_vec_kernel(MergeArgs a, float* __restrict__ m, double* __restrict__ stats_out,
                 int C, int L, int chunks_per_sample) {
    __shared__ SampleNorm s_norm[kMaxDepth];
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x < a.depth) s_norm[threadIdx.x] = sample_norm(a.n[threadIdx.x], sample);
    __syncthreads();

    const int QR = L >> 2;
    const int items = C * QR;
    float acc_s = 0.f, acc_q = 0.f;
#pragma unroll
    for (int it = 0; it < kMgItems; ++it) {
        const int item = (chunk * kMgItems + it) * kMgThreads + threadIdx.x;
        if (item < items) {
            const int c = item / QR;
            const int q = item - c * QR;
            const size_t row = (size_t)sample * C + c;
            float o[4];
            {
                const ChanNorm cn = chan_norm(a.n[0], s_norm[0], c);
                const float4 v = ldg4(a.z[0] + row * L + 4 * q);
                o[0] = apply_norm(cn, v.x); o[1] = apply_norm(cn, v.y);
                o[2] = apply_norm(cn, v.z); o[3] = apply_norm(cn, v.w);
            }
            if (a.depth > 1) {
                const ChanNorm cn = chan_norm(a.n[1], s_norm[1], c);
                const float2 v = __ldg(reinterpret_cast<const float2*>(a.z[1] + row * (L >> 1) + 2 * q));
                const float v0 = apply_norm(cn, v.x), v1 = apply_norm(cn, v.y);
                o[0] += v0; o[1] += v0; o[2] += v1; o[3] += v1;
            }
            for (int d = 2; d < a.depth; ++d) {
                const ChanNorm cn = chan_norm(a.n[d], s_norm[d], c);
                const float v = apply_norm(cn, __ldg(a.z[d] + row * (L >> d) + (q >> (d - 2))));
                o[0] += v; o[1] += v; o[2] += v; o[3] += v;
            }
            *reinterpret_cast<float4*>(m + row * L + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc_s += o[i]; acc_q = fmaf(o[i], o[i], acc_q); }
        }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}

__global__ void __launch_bounds__(kMgThreads)
merge_scalar_kernel(MergeArgs a, float* __restrict__ m, double* __restrict__ stats_out,
                    int C, int L, int chunks_per_sample) {
    __shared__ SampleNorm s_norm[kMaxDepth];
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x < a.depth) s_norm[threadIdx.x] = sample_norm(a.n[threadIdx.x], sample);
    __syncthreads();
    const int items = C * L;
    float acc_s = 0.f, acc_q = 0.f;
    for (int it = 0; it < 4; ++it) {
        const int item = (chunk * 4 + it) * kMgThreads + threadIdx.x;
        if (item < items) {
            const int c = item / L;
            const int t = item - c * L;
            const size_t row = (size_t)sample * C + c;
            float o = 0.f;
            for (int d = 0; d < a.depth; ++d) {
                const ChanNorm cn = chan_norm(a.n[d], s_norm[d], c);
                o += apply_norm(cn, __ldg(a.z[d] + row * (L >> d) + (t >> d)));
            }
            m[row * L + t] = o;
            acc_s += o; acc_q = fmaf(o, o, acc_q);
        }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
int launch_depthwise(const float* x, const NormIn& nin, const float* w5, const float* bias,
                     float* y, double* stats_out, int samples, int C, int Lin, int stride,
                     cudaStream_t st) {
    if (samples <= 0 || C <= 0 || Lin <= 0 || (stride != 1 && stride != 2)) return SDR_ERR_BAD_ARGUMENT;
    const int Lout = (Lin + 4 - 5) / stride + 1;
    const bool vec = (Lout % 4 == 0) && (stride == 1 || Lin == 2 * Lout) &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 == 0);
    if (vec) {
        const long long items = (long long)C * (Lout / 4);
        const int chunks = (int)((items + kDwThreads * kDwItems - 1) / (kDwThreads * kDwItems));
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        if (stride == 1)
            dw5_vec_kernel<1><<<(unsigned)grid, kDwThreads, 0, st>>>(x, nin, w5, bias, y, stats_out, C, Lin, Lout, chunks);
        else
            dw5_vec_kernel<2><<<(unsigned)grid, kDwThreads, 0, st>>>(x, nin, w5, bias, y, stats_out, C, Lin, Lout, chunks);
    } else {
        const long long items = (long long)C * Lout;
        const int chunks = (int)((items + kDwThreads * 4 - 1) / (kDwThreads * 4));
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        dw5_scalar_kernel<<<(unsigned)grid, kDwThreads, 0, st>>>(x, nin, w5, bias, y, stats_out, C, Lin, Lout, stride, chunks);
    }
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}

int launch_merge(const float* const* z, const NormIn* nins, int depth, float* m, double* stats_out,
                 int samples, int C, int L, cudaStream_t st) {
    if (depth < 1 || depth > kMaxDepth) return SDR_ERR_UNSUPPORTED;
    if (samples <= 0 || C <= 0 || L <= 0 || (L % (1 << (depth - 1))) != 0) return SDR_ERR_BAD_ARGUMENT;
    MergeArgs a;
    a.depth = depth;
    bool aligned = reinterpret_cast<uintptr_t>(m) % 16 == 0;
    for (int d = 0; d < kMaxDepth; ++d) {
        a.z[d] = d < depth ? z[d] : nullptr;
        a.n[d] = d < depth ? nins[d] : NormIn{nullptr, nullptr, nullptr, nullptr, 1.0};
        if (d < depth) aligned = aligned && reinterpret_cast<uintptr_t>(z[d]) % 16 == 0;
    }
    // vector path: rows of level 0 are float4-aligned, rows of level 1 float2-aligned
    const bool vec = aligned && (L % 4 == 0);
    if (vec) {
        const long long items = (long long)C * (L / 4);
        const int chunks = (int)((items + kMgThreads * kMgItems - 1) / (kMgThreads * kMgItems));
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        merge_vec_kernel<<<(unsigned)grid, kMgThreads, 0, st>>>(a, m, stats_out, C, L, chunks);
    } else {
        const long long items = (long long)C * L;
        const int chunks = (int)((items + kMgThreads * 4 - 1) / (kMgThreads * 4));
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        merge_scalar_kernel<<<(unsigned)grid, kMgThreads, 0, st>>>(a, m, stats_out, C, L, chunks);
    }
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}

}  // namespace sdr
