// HBM-bound elementwise kernels: BCE-with-logits (fused forward + backward), multi-tensor Adam,
// activation backward.  Reference call sites: nn.BCEWithLogitsLoss (network_tests.py:248,304-306,313;
// SIMNN.py:257), torch.optim.Adam (network_tests.py:253-254,308,315; SIMNN.py:258-259,316,331).
#include "common.cuh"

namespace {

// loss_i = max(x,0) - x*y + log1p(exp(-|x|));   dL/dx_i = (sigmoid(x) - y) * gscale
__global__ void __launch_bounds__(1024) bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ y, float y_const,
                                                           long long n, float inv_n, const float* __restrict__ gscale_ptr, float gscale,
                                                           float* __restrict__ loss, int accumulate, float* __restrict__ dx) {
    __shared__ double red[32];
    double acc = 0.0;
    const float gs = gscale_ptr ? gscale * gscale_ptr[0] : gscale;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float xi = x[i], yi = y ? y[i] : y_const;
        if (loss) acc += (double)(fmaxf(xi, 0.f) - xi * yi + log1pf(expf(-fabsf(xi))));
        if (dx) dx[i] = (1.f / (1.f + expf(-xi)) - yi) * gs;
    }
    if (!loss) return;
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) loss[0] = (accumulate ? loss[0] : 0.f) + (float)(v * (double)inv_n);
    }
}

__global__ void act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dz, long long n, int act) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dz[i] = dy[i] * mmg_act_grad(y[i], act);
}

__global__ void fill_scalar_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    const float v = src[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = v;
}

__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, long long n, float* __restrict__ out, int accumulate) {
    __shared__ double red[32];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += (double)x[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = red[threadIdx.x];
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + (float)v;
    }
}

constexpr int ADAM_MAX_TENSORS = 48;
struct AdamTable {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    long long n[ADAM_MAX_TENSORS];
};

// torch/optim/adam.py single-tensor update:
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); denom = sqrt(v)/sqrt(bc2) + eps; p -= (lr/bc1) * m/denom
__global__ void __launch_bounds__(256) adam_multi_tensor_kernel(AdamTable tab, float beta1, float beta2, float step_size, float bc2_sqrt,
                                                                 float eps, float grad_scale) {
    const int t = blockIdx.y;
    const long long n = tab.n[t];
    float* __restrict__ p = tab.p[t];
    const float* __restrict__ g = tab.g[t];
    float* __restrict__ m = tab.m[t];
    float* __restrict__ v = tab.v[t];
    const float w = 1.f - beta1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
    const long long n4 = vec ? n / 4 : 0;
    auto upd = [&](float& pi, float gi, float& mi, float& vi) {
        gi *= grad_scale;
        mi = (w < 0.5f) ? mi + w * (gi - mi) : gi - (gi - mi) * (1.f - w);      // at::lerp
        vi = vi * beta2 + (1.f - beta2) * gi * gi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi = pi - step_size * (mi / denom);
    };
    for (long long i = i0; i < n4; i += stride) {
        float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        const float4 g4 = reinterpret_cast<const float4*>(g)[i];
        upd(p4.x, g4.x, m4.x, v4.x); upd(p4.y, g4.y, m4.y, v4.y); upd(p4.z, g4.z, m4.z, v4.z); upd(p4.w, g4.w, m4.w, v4.w);
        reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
    }
    for (long long i = n4 * 4 + i0; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
}

// Device-resident hyper-parameters (CUDA-graph replay): hyper = [lr, beta1, beta2, eps], step_dev = number of
// updates already applied; this launch applies update number *step_dev + 1 (the counter is bumped by adam_step_inc_kernel).
__global__ void __launch_bounds__(256) adam_multi_tensor_dev_kernel(AdamTable tab, const float* __restrict__ hyper, const long long* __restrict__ step_dev,
                                                                     float grad_scale) {
    const float lr = hyper[0], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3];
    const double step = (double)(step_dev[0] + 1);
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const int t = blockIdx.y;
    const long long n = tab.n[t];
    float* __restrict__ p = tab.p[t];
    const float* __restrict__ g = tab.g[t];
    float* __restrict__ m = tab.m[t];
    float* __restrict__ v = tab.v[t];
    const float w = 1.f - beta1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // 16-byte accesses when the four tensors allow it (the scalar form moved fc1's 198 MB at 2.6 TB/s); same arithmetic per element either way
    const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
    const long long n4 = vec ? n / 4 : 0;
    auto upd = [&](float& pi, float gi, float& mi, float& vi) {
        gi *= grad_scale;
        mi = (w < 0.5f) ? mi + w * (gi - mi) : gi - (gi - mi) * (1.f - w);      // at::lerp
        vi = vi * beta2 + (1.f - beta2) * gi * gi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi = pi - step_size * (mi / denom);
    };
    for (long long i = i0; i < n4; i += stride) {
        float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        const float4 g4 = reinterpret_cast<const float4*>(g)[i];
        upd(p4.x, g4.x, m4.x, v4.x); upd(p4.y, g4.y, m4.y, v4.y); upd(p4.z, g4.z, m4.z, v4.z); upd(p4.w, g4.w, m4.w, v4.w);
        reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
    }
    for (long long i = n4 * 4 + i0; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
}
__global__ void adam_step_inc_kernel(long long* step_dev) { step_dev[0] += 1; }

}  // namespace

extern "C" {

// mean-reduced BCE-with-logits.  loss (1 float, device; may be null) gets  [loss +]= mean_i(l_i);
// dlogits (may be null) gets (sigmoid(x)-y) * gscale * (*gscale_dev if given)  -- pass gscale = upstream/n.
int mmg_bce_logits_f32(const float* logits, const float* targets, float target_const, int64_t n, float* loss, int accumulate,
                       float* dlogits, float gscale, const float* gscale_dev, void* stream) {
    MMG_REQUIRE(n > 0 && logits, MMG_EINVAL, "bce: empty input");
    bce_logits_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, targets, target_const, n, 1.f / (float)n, gscale_dev, gscale, loss,
                                                           accumulate, dlogits);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// dst[0..n) = *src  (e.g. logits initialised with the fc bias before the fused conv2+fc kernel accumulates into them)
int mmg_fill_scalar_f32(float* dst, const float* src_dev, int64_t n, void* stream) {
    MMG_REQUIRE(n >= 0 && (n == 0 || (dst && src_dev)), MMG_EINVAL, "fill_scalar: bad arguments");
    if (n == 0) return MMG_OK;
    fill_scalar_kernel<<<mmg_grid(n, 256, 2), 256, 0, (cudaStream_t)stream>>>(dst, src_dev, n);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// stream-ordered zero fill (gradient buffers)
int mmg_zero(void* dst, size_t bytes, void* stream) {
    MMG_REQUIRE(bytes == 0 || dst, MMG_EINVAL, "zero: null pointer");
    if (bytes) MMG_CUDA(cudaMemsetAsync(dst, 0, bytes, (cudaStream_t)stream));
    return MMG_OK;
}

// out[0] (+)= sum(x)   (fc bias gradient = sum of dlogits)
int mmg_sum_f32(const float* x, int64_t n, float* out, int accumulate, void* stream) {
    MMG_REQUIRE(n >= 0 && out, MMG_EINVAL, "sum: bad arguments");
    sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, out, accumulate);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

int mmg_act_bwd_f32(const float* y, const float* dy, float* dz, int64_t n, int act, void* stream) {
    MMG_REQUIRE(n >= 0, MMG_EINVAL, "act_bwd: negative size");
    if (n == 0) return MMG_OK;
    act_bwd_kernel<<<mmg_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(y, dy, dz, n, act);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// One launch per <= 48 tensors.  ptrs: 4*n_tensors device pointers laid out [p0..][g0..][m0..][v0..] in HOST memory.
int mmg_adam_multi_tensor_f32(int n_tensors, void* const* ptrs, const int64_t* sizes, float lr, float beta1, float beta2, float eps,
                              int64_t step, float grad_scale, void* stream) {
    MMG_REQUIRE(n_tensors >= 0 && step >= 1, MMG_EINVAL, "adam: bad n_tensors/step");
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    for (int t0 = 0; t0 < n_tensors; t0 += ADAM_MAX_TENSORS) {
        const int nt = n_tensors - t0 < ADAM_MAX_TENSORS ? n_tensors - t0 : ADAM_MAX_TENSORS;
        AdamTable tab;
        long long maxn = 0;
        for (int i = 0; i < nt; ++i) {
            tab.p[i] = (float*)ptrs[t0 + i];
            tab.g[i] = (const float*)ptrs[n_tensors + t0 + i];
            tab.m[i] = (float*)ptrs[2 * n_tensors + t0 + i];
            tab.v[i] = (float*)ptrs[3 * n_tensors + t0 + i];
            tab.n[i] = sizes[t0 + i];
            MMG_REQUIRE(tab.n[i] >= 0 && (tab.n[i] == 0 || (tab.p[i] && tab.g[i] && tab.m[i] && tab.v[i])), MMG_EINVAL, "adam: null tensor %d", t0 + i);
            if (tab.n[i] > maxn) maxn = tab.n[i];
        }
        if (maxn == 0) continue;
        dim3 grid(mmg_grid((maxn + 3) / 4, 256, 4), nt);
        adam_multi_tensor_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tab, beta1, beta2, step_size, bc2_sqrt, eps, grad_scale);
        MMG_LAUNCH_CHECK();
    }
    return MMG_OK;
}

// Same update with the hyper-parameters and the step counter in DEVICE memory, so that the launch can be replayed from a
// CUDA graph: hyper_dev = [lr, beta1, beta2, eps] (fp32), step_dev = updates applied so far (int64, incremented here).
int mmg_adam_multi_tensor_dev_f32(int n_tensors, void* const* ptrs, const int64_t* sizes, const float* hyper_dev, int64_t* step_dev,
                                  float grad_scale, void* stream) {
    MMG_REQUIRE(n_tensors >= 0 && n_tensors <= ADAM_MAX_TENSORS && hyper_dev && step_dev, MMG_EINVAL, "adam_dev: bad arguments (at most %d tensors)", ADAM_MAX_TENSORS);
    if (n_tensors == 0) return MMG_OK;
    AdamTable tab;
    long long maxn = 0;
    for (int i = 0; i < n_tensors; ++i) {
        tab.p[i] = (float*)ptrs[i];
        tab.g[i] = (const float*)ptrs[n_tensors + i];
        tab.m[i] = (float*)ptrs[2 * n_tensors + i];
        tab.v[i] = (float*)ptrs[3 * n_tensors + i];
        tab.n[i] = sizes[i];
        MMG_REQUIRE(tab.n[i] >= 0 && (tab.n[i] == 0 || (tab.p[i] && tab.g[i] && tab.m[i] && tab.v[i])), MMG_EINVAL, "adam_dev: null tensor %d", i);
        if (tab.n[i] > maxn) maxn = tab.n[i];
    }
    if (maxn) {
        dim3 grid(mmg_grid(maxn, 256, 4), n_tensors);
        adam_multi_tensor_dev_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tab, hyper_dev, (const long long*)step_dev, grad_scale);
        MMG_LAUNCH_CHECK();
    }
    adam_step_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)step_dev);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
