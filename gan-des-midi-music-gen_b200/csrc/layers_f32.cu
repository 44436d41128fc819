// fp32 SIMT layer kernels: the full-precision path (fp32 operands, fp32 FMA accumulation) used for
// reference-tolerance parity (losses rel 1e-5) and for the GAN-DES family.
// All tensors are contiguous: activations NCHW / (rows, features), Linear weight (out,in),
// Conv2d weight (Co,Ci,kh,kw), ConvTranspose2d weight (Ci,Co,kh,kw) (== the Conv2d weight of the
// convolution whose data-gradient it is).
// Reference call sites: nn.Linear/BatchNorm1d/Sigmoid (network_tests.py:75-80), Conv2d+LeakyReLU
// (:150-158), GAN-DES ConvTranspose2d/BatchNorm2d/ReLU (SIMNN.py:70-110), Conv2d/ReLU/MaxPool2d/Linear
// (SIMNN.py:123-141).
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// SGEMM, generic strides:  C[M,N] (+)= act( sum_k A(i,k) B(k,j) + bias[j] )
// ------------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, long long sa_i, long long sa_k,
                                                     const float* __restrict__ B, long long sb_k, long long sb_j,
                                                     float* __restrict__ C, int M, int N, int K, const float* __restrict__ bias, int act,
                                                     int k_per_split, int atomic_out, int accumulate) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_lo = blockIdx.z * k_per_split, k_hi = min(K, k_lo + k_per_split);
    float acc[4][4] = {};
    for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int e = tid + l * 256;
            int i, k;
            if (sa_k == 1) { k = e & 15; i = e >> 4; } else { i = e & 63; k = e >> 6; }
            const int gi = m0 + i, gk = k0 + k;
            As[k][i] = (gi < M && gk < k_hi) ? A[gi * sa_i + gk * sa_k] : 0.f;
            int j, kb;
            if (sb_k == 1) { kb = e & 15; j = e >> 4; } else { j = e & 63; kb = e >> 6; }
            const int gj = n0 + j, gkb = k0 + kb;
            Bs[kb][j] = (gj < N && gkb < k_hi) ? B[gkb * sb_k + gj * sb_j] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[k][ty * 4 + r];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = Bs[k][tx * 4 + c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int gi = m0 + ty * 4 + r;
        if (gi >= M) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gj = n0 + tx * 4 + c;
            if (gj >= N) continue;
            float v = acc[r][c];
            float* dst = C + (size_t)gi * N + gj;
            if (atomic_out) {
                if (bias && blockIdx.z == 0) v += bias[gj];
                atomicAdd(dst, v);
            } else {
                if (bias) v += bias[gj];
                v = mmg_act(v, act);
                *dst = accumulate ? *dst + v : v;
            }
        }
    }
}

// y[i][j] = act(y[i][j] + bias[j])  (epilogue of a split-K product whose partial sums were accumulated with atomics)
__global__ void bias_act_inplace_kernel(float* __restrict__ y, const float* __restrict__ bias, long long total, int N, int act) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        y[i] = mmg_act(y[i] + (bias ? bias[i % N] : 0.f), act);
}

int launch_sgemm(const float* A, long long sa_i, long long sa_k, const float* B, long long sb_k, long long sb_j, float* C, int M, int N,
                 int K, const float* bias, int act, int accumulate, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return MMG_OK;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
    int splits = 1;
    const long long tiles = (long long)grid.x * grid.y;
    if (act != MMG_ACT_NONE && !accumulate && K >= 8192 && tiles < MMG_NUM_SMS / 4) {
        // a handful of output tiles over a very long K (GAN-DES fc1: 30 x 128 over 55 296): split K, then bias + activation in place
        int rc = launch_sgemm(A, sa_i, sa_k, B, sb_k, sb_j, C, M, N, K, nullptr, MMG_ACT_NONE, 0, stream);
        if (rc) return rc;
        const long long total = (long long)M * N;
        bias_act_inplace_kernel<<<mmg_grid(total, 256), 256, 0, stream>>>(C, bias, total, N, act);
        MMG_LAUNCH_CHECK();
        return MMG_OK;
    }
    if (act == MMG_ACT_NONE && K >= 2048 && tiles < 2 * MMG_NUM_SMS) {        // split-K for skinny outputs (weight grads)
        splits = (int)((4LL * MMG_NUM_SMS + tiles - 1) / tiles);
        const int max_splits = (K + 255) / 256;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    int k_per_split = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    if (k_per_split < BK) k_per_split = BK;
    splits = K > 0 ? (K + k_per_split - 1) / k_per_split : 1;
    grid.z = splits;
    const int atomic_out = splits > 1;
    if (atomic_out && !accumulate) MMG_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, stream));
    sgemm_kernel<<<grid, 256, 0, stream>>>(A, sa_i, sa_k, B, sb_k, sb_j, C, M, N, K, bias, act, k_per_split, atomic_out, accumulate);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// out[j] (+)= sum_i a[i*N + j]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ a, float* __restrict__ out, long long M, int N, long long rows_per_block) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + cx;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float acc = 0.f;
    if (j < N)
        for (long long i = r0 + ry; i < r1; i += 8) acc += a[i * N + j];
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && j < N) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][cx];
        atomicAdd(&out[j], s);
    }
}

int launch_colsum(const float* a, float* out, long long M, int N, int accumulate, cudaStream_t stream) {
    if (N <= 0) return MMG_OK;
    if (!accumulate) MMG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, stream));
    if (M <= 0) return MMG_OK;
    const int gx = (N + 31) / 32;
    long long gy = (2LL * MMG_NUM_SMS + gx - 1) / gx;
    if (gy > (M + 63) / 64) gy = (M + 63) / 64;
    if (gy < 1) gy = 1;
    const long long rpb = (M + gy - 1) / gy;
    colsum_kernel<<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(a, out, M, N, rpb);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// ------------------------------------------------------------------------------------------------
// BatchNorm over (N, C, HW):  per-channel statistics in double, normalise + activation fused
// ------------------------------------------------------------------------------------------------
// stats[c] += sum z, stats[C+c] += sum z^2 ; also used for the backward sums (sum g, sum g*xhat)
template <bool BWD>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ z, const float* __restrict__ dy, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ mean,
                                                         const float* __restrict__ invstd, double* __restrict__ stats, long long N, int C,
                                                         int HW, int act, long long n_per_block) {
    const long long n0 = (long long)blockIdx.y * n_per_block, n1 = min(N, n0 + n_per_block);
    double s0 = 0.0, s1 = 0.0;
    int c;
    if (HW == 1) {      // (N, C) row-major: 32 channels x 8 row lanes per block, coalesced along C
        c = blockIdx.x * 32 + (threadIdx.x & 31);
        if (c < C) {
            float mu = 0.f, is = 0.f, ga = 0.f, be = 0.f;
            if (BWD) { mu = mean[c]; is = invstd[c]; ga = gamma[c]; be = beta[c]; }
            for (long long n = n0 + (threadIdx.x >> 5); n < n1; n += 8) {
                const float v = z[n * C + c];
                if (BWD) {
                    const float xh = (v - mu) * is;
                    const float g = dy[n * C + c] * mmg_act_grad(mmg_act(xh * ga + be, act), act);
                    s0 += g; s1 += (double)g * xh;
                } else { s0 += v; s1 += (double)v * v; }
            }
        }
        __shared__ double r0[8][33], r1[8][33];
        r0[threadIdx.x >> 5][threadIdx.x & 31] = s0;
        r1[threadIdx.x >> 5][threadIdx.x & 31] = s1;
        __syncthreads();
        if (threadIdx.x < 32 && c < C) {
            double a = 0, b = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) { a += r0[r][threadIdx.x]; b += r1[r][threadIdx.x]; }
            atomicAdd(&stats[c], a);
            atomicAdd(&stats[C + c], b);
        }
    } else {            // (N, C, HW): one channel per block.x, coalesced along HW
        c = blockIdx.x;
        float mu = 0.f, is = 0.f, ga = 0.f, be = 0.f;
        if (BWD) { mu = mean[c]; is = invstd[c]; ga = gamma[c]; be = beta[c]; }
        const long long cnt = (n1 - n0) * HW;
        for (long long e = threadIdx.x; e < cnt; e += 256) {
            const long long n = n0 + e / HW;
            const int hw = (int)(e % HW);
            const size_t idx = ((size_t)n * C + c) * HW + hw;
            const float v = z[idx];
            if (BWD) {
                const float xh = (v - mu) * is;
                const float g = dy[idx] * mmg_act_grad(mmg_act(xh * ga + be, act), act);
                s0 += g; s1 += (double)g * xh;
            } else { s0 += v; s1 += (double)v * v; }
        }
        __shared__ double q0[8], q1[8];
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if ((threadIdx.x & 31) == 0) { q0[threadIdx.x >> 5] = s0; q1[threadIdx.x >> 5] = s1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0, b = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) { a += q0[r]; b += q1[r]; }
            atomicAdd(&stats[c], a);
            atomicAdd(&stats[C + c], b);
        }
    }
}

// mean / invstd from the sums; running stats: momentum update with the UNBIASED variance
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, double count, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* __restrict__ run_mean, float* __restrict__ run_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mu = stats[c] / count;
    double var = stats[C + c] / count - mu * mu;
    if (var < 0) var = 0;
    mean[c] = (float)mu;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (run_mean) {
        const float unb = (float)(var * (count / (count - 1.0)));
        run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)mu;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * unb;
    }
}

// y = act((z-mean)*invstd*gamma + beta).  eval mode passes running stats: invstd_from_var=1 -> invstd = rsqrt(var+eps)
__global__ void bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ mean, const float* __restrict__ invstd_or_var, int invstd_from_var, float eps,
                                float* __restrict__ y, long long total, int C, int HW, int act) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / HW) % C);
        const float is = invstd_from_var ? 1.f / sqrtf(invstd_or_var[c] + eps) : invstd_or_var[c];
        y[i] = mmg_act((z[i] - mean[c]) * is * gamma[c] + beta[c], act);
    }
}

// dz = gamma*invstd*(g - sum_g/n - xhat*sum_gx/n), g = dy*act'(y);  dgamma = sum_gx, dbeta = sum_g
__global__ void bn_bwd_apply_kernel(const float* __restrict__ z, const float* __restrict__ dy, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const double* __restrict__ sums, double count, float* __restrict__ dz, long long total, int C, int HW, int act) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / HW) % C);
        const float xh = (z[i] - mean[c]) * invstd[c];
        const float g = dy[i] * mmg_act_grad(mmg_act(xh * gamma[c] + beta[c], act), act);
        const float sg = (float)(sums[c] / count), sgx = (float)(sums[C + c] / count);
        dz[i] = gamma[c] * invstd[c] * (g - sg - xh * sgx);
    }
}

__global__ void bn_bwd_params_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float dg = (float)sums[C + c], db = (float)sums[c];
    dgamma[c] = accumulate ? dgamma[c] + dg : dg;
    dbeta[c] = accumulate ? dbeta[c] + db : db;
}

template <bool BWD>
int launch_bn_reduce(const float* z, const float* dy, const float* gamma, const float* beta, const float* mean, const float* invstd,
                     double* stats, long long N, int C, int HW, int act, cudaStream_t stream) {
    MMG_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, stream));
    const int gx = HW == 1 ? (C + 31) / 32 : C;
    long long gy = (4LL * MMG_NUM_SMS + gx - 1) / gx;
    const long long min_rows = HW == 1 ? 64 : 1;
    if (gy > (N + min_rows - 1) / min_rows) gy = (N + min_rows - 1) / min_rows;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const long long npb = (N + gy - 1) / gy;
    bn_reduce_kernel<BWD><<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(z, dy, gamma, beta, mean, invstd, stats, N, C, HW, act, npb);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// ------------------------------------------------------------------------------------------------
// direct Conv2d forward / data-gradient / weight-gradient (generic kernel size, stride, padding)
// ------------------------------------------------------------------------------------------------
struct ConvDims { int N, Ci, H, W, Co, kh, kw, stride, pad, OH, OW; };

__global__ void __launch_bounds__(256) conv2d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                          float* __restrict__ y, ConvDims d, int act) {
    const long long total = (long long)d.N * d.Co * d.OH * d.OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % d.OW), oy = (int)((i / d.OW) % d.OH);
        const int co = (int)((i / ((long long)d.OW * d.OH)) % d.Co);
        const long long n = i / ((long long)d.OW * d.OH * d.Co);
        float acc = b ? b[co] : 0.f;
        const float* wp = w + (size_t)co * d.Ci * d.kh * d.kw;
        const int iy0 = oy * d.stride - d.pad, ix0 = ox * d.stride - d.pad;
        for (int ci = 0; ci < d.Ci; ++ci) {
            const float* xp = x + ((size_t)n * d.Ci + ci) * d.H * d.W;
            for (int ky = 0; ky < d.kh; ++ky) {
                const int iy = iy0 + ky;
                if (iy < 0 || iy >= d.H) continue;
                for (int kx = 0; kx < d.kw; ++kx) {
                    const int ix = ix0 + kx;
                    if (ix < 0 || ix >= d.W) continue;
                    acc = fmaf(xp[iy * d.W + ix], wp[(ci * d.kh + ky) * d.kw + kx], acc);
                }
            }
        }
        y[i] = mmg_act(acc, act);
    }
}

// dx[n,ci,iy,ix] = sum_{co,ky,kx} dy[n,co,oy,ox] * w[co,ci,ky,kx]  with iy = oy*stride - pad + ky
// (also the ConvTranspose2d forward; bias/act are applied on that use)
__global__ void __launch_bounds__(256) conv2d_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ b,
                                                               float* __restrict__ dx, ConvDims d, int act) {
    const long long total = (long long)d.N * d.Ci * d.H * d.W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % d.W), iy = (int)((i / d.W) % d.H);
        const int ci = (int)((i / ((long long)d.W * d.H)) % d.Ci);
        const long long n = i / ((long long)d.W * d.H * d.Ci);
        float acc = b ? b[ci] : 0.f;
        for (int ky = 0; ky < d.kh; ++ky) {
            const int ty = iy + d.pad - ky;
            if (ty < 0 || ty % d.stride) continue;
            const int oy = ty / d.stride;
            if (oy >= d.OH) continue;
            for (int kx = 0; kx < d.kw; ++kx) {
                const int tx = ix + d.pad - kx;
                if (tx < 0 || tx % d.stride) continue;
                const int ox = tx / d.stride;
                if (ox >= d.OW) continue;
                const float* dyp = dy + (size_t)n * d.Co * d.OH * d.OW + (size_t)oy * d.OW + ox;
                const float* wp = w + ((size_t)ci * d.kh + ky) * d.kw + kx;
                for (int co = 0; co < d.Co; ++co)
                    acc = fmaf(dyp[(size_t)co * d.OH * d.OW], wp[(size_t)co * d.Ci * d.kh * d.kw], acc);
            }
        }
        dx[i] = mmg_act(acc, act);
    }
}

// dw[co,ci,ky,kx] += sum_{n,oy,ox} dy[n,co,oy,ox] * x[n,ci,iy,ix];  db[co] += sum dy   (block = (co,ci) x sample slice)
constexpr int WG_MAX_TAPS = 25;
__global__ void __launch_bounds__(256) conv2d_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                                                 float* __restrict__ db, ConvDims d, long long n_per_block) {
    const int co = blockIdx.x / d.Ci, ci = blockIdx.x % d.Ci;
    const long long n0 = (long long)blockIdx.y * n_per_block, n1 = min((long long)d.N, n0 + n_per_block);
    const int taps = d.kh * d.kw, P = d.OH * d.OW;
    float acc[WG_MAX_TAPS];
#pragma unroll
    for (int t = 0; t < WG_MAX_TAPS; ++t) acc[t] = 0.f;
    float bsum = 0.f;
    const long long cnt = (n1 - n0) * P;
    for (long long e = threadIdx.x; e < cnt; e += 256) {
        const long long n = n0 + e / P;
        const int pos = (int)(e % P), oy = pos / d.OW, ox = pos % d.OW;
        const float g = dy[((size_t)n * d.Co + co) * P + pos];
        bsum += g;
        const float* xp = x + ((size_t)n * d.Ci + ci) * d.H * d.W;
        const int iy0 = oy * d.stride - d.pad, ix0 = ox * d.stride - d.pad;
#pragma unroll
        for (int t = 0; t < WG_MAX_TAPS; ++t) {
            if (t < taps) {
                const int iy = iy0 + t / d.kw, ix = ix0 + t % d.kw;
                if (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) acc[t] = fmaf(g, xp[iy * d.W + ix], acc[t]);
            }
        }
    }
    __shared__ float red[8];
    auto block_sum = [&](float v) -> float {
        v = warp_sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        float s = 0.f;
        if (threadIdx.x == 0)
#pragma unroll
            for (int r = 0; r < 8; ++r) s += red[r];
        return s;
    };
#pragma unroll
    for (int t = 0; t < WG_MAX_TAPS; ++t) {
        if (t < taps) {
            const float s = block_sum(acc[t]);
            if (threadIdx.x == 0) atomicAdd(&dw[((size_t)co * d.Ci + ci) * taps + t], s);
        }
    }
    if (db && ci == 0) {
        const float s = block_sum(bsum);
        if (threadIdx.x == 0) atomicAdd(&db[co], s);
    }
}

// MaxPool2d(kernel 2, stride 2, no padding), floor mode; idx = argmax within the 2x2 window (0..3)
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ idx, long long NC, int H, int W, int OH, int OW) {
    const long long total = NC * OH * OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
        const long long nc = i / ((long long)OW * OH);
        const float* p = x + (size_t)nc * H * W + (size_t)(2 * oy) * W + 2 * ox;
        float best = p[0];
        int bi = 0;
        const float v1 = p[1], v2 = p[W], v3 = p[W + 1];
        if (v1 > best || v1 != v1) { best = v1; bi = 1; }
        if (v2 > best || v2 != v2) { best = v2; bi = 2; }
        if (v3 > best || v3 != v3) { best = v3; bi = 3; }
        y[i] = best;
        if (idx) idx[i] = (uint8_t)bi;
    }
}
__global__ void maxpool2_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx, float* __restrict__ dx, long long NC, int H, int W, int OH, int OW) {
    const long long total = NC * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % W), iy = (int)((i / W) % H);
        const long long nc = i / ((long long)W * H);
        const int oy = iy >> 1, ox = ix >> 1;
        float v = 0.f;
        if (oy < OH && ox < OW) {
            const size_t o = (size_t)nc * OH * OW + (size_t)oy * OW + ox;
            if (idx[o] == ((iy & 1) * 2 + (ix & 1))) v = dy[o];
        }
        dx[i] = v;
    }
}

int fill_conv_dims(ConvDims& d, int N, int Ci, int H, int W, int Co, int kh, int kw, int stride, int pad) {
    MMG_REQUIRE(N >= 0 && Ci > 0 && Co > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, MMG_EINVAL, "conv: bad dims");
    d = {N, Ci, H, W, Co, kh, kw, stride, pad, (H + 2 * pad - kh) / stride + 1, (W + 2 * pad - kw) / stride + 1};
    MMG_REQUIRE(d.OH > 0 && d.OW > 0, MMG_EINVAL, "conv: empty output");
    return MMG_OK;
}

}  // namespace

extern "C" {

// y[M,N] = act(x[M,K] . w[N,K]^T + b)
int mmg_linear_fwd_f32(const float* x, const float* w, const float* b, float* y, int64_t M, int64_t N, int64_t K, int act, void* stream) {
    MMG_REQUIRE(M >= 0 && N > 0 && K > 0 && M < (1LL << 31), MMG_EINVAL, "linear_fwd: bad dims");
    return launch_sgemm(x, K, 1, w, 1, K, y, (int)M, (int)N, (int)K, b, act, 0, (cudaStream_t)stream);
}

// dx[M,K] = dy.w (if dx);  dw[N,K] (+)= dy^T.x (if dw);  db[N] (+)= colsum(dy) (if db)
int mmg_linear_bwd_f32(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int64_t M, int64_t N, int64_t K,
                       int accumulate, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMG_REQUIRE(M >= 0 && N > 0 && K > 0 && M < (1LL << 31), MMG_EINVAL, "linear_bwd: bad dims");
    int rc;
    if (dx && (rc = launch_sgemm(dy, N, 1, w, K, 1, dx, (int)M, (int)K, (int)N, nullptr, MMG_ACT_NONE, 0, stream))) return rc;
    if (dw && (rc = launch_sgemm(dy, 1, N, x, K, 1, dw, (int)N, (int)K, (int)M, nullptr, MMG_ACT_NONE, accumulate, stream))) return rc;
    if (db && (rc = launch_colsum(dy, db, M, (int)N, accumulate, stream))) return rc;
    return MMG_OK;
}

size_t mmg_bn_workspace_bytes(int64_t C) { return sizeof(double) * 2 * (size_t)C; }

// training-mode BatchNorm over (N, C, HW) fused with activation; saves mean / invstd for the backward.
int mmg_bn_fwd_train_f32(const float* z, const float* gamma, const float* beta, float* run_mean, float* run_var, float* y, float* save_mean,
                         float* save_invstd, int64_t N, int64_t C, int64_t HW, float momentum, float eps, int act, void* workspace,
                         size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMG_REQUIRE(N > 0 && C > 0 && HW > 0, MMG_EINVAL, "bn: bad dims");
    MMG_REQUIRE(N * HW > 1, MMG_EINVAL, "Expected more than 1 value per channel when training");
    MMG_REQUIRE(workspace && ws_bytes >= mmg_bn_workspace_bytes(C), MMG_EWORKSPACE, "bn: workspace too small");
    double* stats = (double*)workspace;
    int rc = launch_bn_reduce<false>(z, nullptr, nullptr, nullptr, nullptr, nullptr, stats, N, (int)C, (int)HW, act, stream);
    if (rc) return rc;
    bn_finalize_kernel<<<((int)C + 127) / 128, 128, 0, stream>>>(stats, (int)C, (double)(N * HW), eps, momentum, save_mean, save_invstd, run_mean, run_var);
    MMG_LAUNCH_CHECK();
    const long long total = N * C * HW;
    bn_apply_kernel<<<mmg_grid(total, 256), 256, 0, stream>>>(z, gamma, beta, save_mean, save_invstd, 0, eps, y, total, (int)C, (int)HW, act);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

int mmg_bn_fwd_eval_f32(const float* z, const float* gamma, const float* beta, const float* run_mean, const float* run_var, float* y, int64_t N,
                        int64_t C, int64_t HW, float eps, int act, void* stream) {
    MMG_REQUIRE(N >= 0 && C > 0 && HW > 0, MMG_EINVAL, "bn: bad dims");
    const long long total = N * C * HW;
    if (total == 0) return MMG_OK;
    bn_apply_kernel<<<mmg_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(z, gamma, beta, run_mean, run_var, 1, eps, y, total, (int)C, (int)HW, act);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// backward of act(BN_train(z)): dz, dgamma (+)=, dbeta (+)=
int mmg_bn_bwd_f32(const float* z, const float* dy, const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                   float* dz, float* dgamma, float* dbeta, int64_t N, int64_t C, int64_t HW, int act, int accumulate, void* workspace,
                   size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMG_REQUIRE(N > 0 && C > 0 && HW > 0, MMG_EINVAL, "bn_bwd: bad dims");
    MMG_REQUIRE(workspace && ws_bytes >= mmg_bn_workspace_bytes(C), MMG_EWORKSPACE, "bn_bwd: workspace too small");
    double* sums = (double*)workspace;
    int rc = launch_bn_reduce<true>(z, dy, gamma, beta, save_mean, save_invstd, sums, N, (int)C, (int)HW, act, stream);
    if (rc) return rc;
    const long long total = N * C * HW;
    if (dz) {
        bn_bwd_apply_kernel<<<mmg_grid(total, 256), 256, 0, stream>>>(z, dy, gamma, beta, save_mean, save_invstd, sums, (double)(N * HW), dz, total, (int)C, (int)HW, act);
        MMG_LAUNCH_CHECK();
    }
    if (dgamma && dbeta) {
        bn_bwd_params_kernel<<<((int)C + 127) / 128, 128, 0, stream>>>(sums, (int)C, dgamma, dbeta, accumulate);
        MMG_LAUNCH_CHECK();
    }
    return MMG_OK;
}

int mmg_conv2d_fwd_f32(const float* x, const float* w, const float* b, float* y, int N, int Ci, int H, int W, int Co, int kh, int kw, int stride,
                       int pad, int act, void* stream) {
    ConvDims d;
    int rc = fill_conv_dims(d, N, Ci, H, W, Co, kh, kw, stride, pad);
    if (rc) return rc;
    const long long total = (long long)N * Co * d.OH * d.OW;
    if (total == 0) return MMG_OK;
    conv2d_fwd_kernel<<<mmg_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(x, w, b, y, d, act);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// data gradient of the conv described by (Ci,H,W)->(Co,OH,OW); b/act are used when this runs as a ConvTranspose2d forward
int mmg_conv2d_bwd_data_f32(const float* dy, const float* w, const float* b, float* dx, int N, int Ci, int H, int W, int Co, int kh, int kw,
                            int stride, int pad, int act, void* stream) {
    ConvDims d;
    int rc = fill_conv_dims(d, N, Ci, H, W, Co, kh, kw, stride, pad);
    if (rc) return rc;
    const long long total = (long long)N * Ci * H * W;
    if (total == 0) return MMG_OK;
    conv2d_bwd_data_kernel<<<mmg_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, w, b, dx, d, act);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

int mmg_conv2d_bwd_weight_f32(const float* x, const float* dy, float* dw, float* db, int N, int Ci, int H, int W, int Co, int kh, int kw,
                              int stride, int pad, int accumulate, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ConvDims d;
    int rc = fill_conv_dims(d, N, Ci, H, W, Co, kh, kw, stride, pad);
    if (rc) return rc;
    MMG_REQUIRE(kh * kw <= WG_MAX_TAPS, MMG_EUNSUPPORTED, "conv wgrad: kernel larger than 5x5");
    if (!accumulate) {
        MMG_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Co * Ci * kh * kw, stream));
        if (db) MMG_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Co, stream));
    }
    if (N == 0) return MMG_OK;
    const int gx = Co * Ci;
    long long gy = (4LL * MMG_NUM_SMS + gx - 1) / gx;
    if (gy > N) gy = N;
    if (gy < 1) gy = 1;
    const long long npb = (N + gy - 1) / gy;
    gy = (N + npb - 1) / npb;
    conv2d_bwd_weight_kernel<<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(x, dy, dw, db, d, npb);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

int mmg_maxpool2_fwd_f32(const float* x, float* y, uint8_t* idx, int64_t NC, int H, int W, void* stream) {
    const int OH = H / 2, OW = W / 2;
    MMG_REQUIRE(NC >= 0 && OH > 0 && OW > 0, MMG_EINVAL, "maxpool: bad dims");
    const long long total = NC * OH * OW;
    if (total == 0) return MMG_OK;
    maxpool2_fwd_kernel<<<mmg_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, idx, NC, H, W, OH, OW);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

int mmg_maxpool2_bwd_f32(const float* dy, const uint8_t* idx, float* dx, int64_t NC, int H, int W, void* stream) {
    const int OH = H / 2, OW = W / 2;
    MMG_REQUIRE(NC >= 0 && OH > 0 && OW > 0, MMG_EINVAL, "maxpool: bad dims");
    const long long total = NC * H * W;
    if (total == 0) return MMG_OK;
    maxpool2_bwd_kernel<<<mmg_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, idx, dx, NC, H, W, OH, OW);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
