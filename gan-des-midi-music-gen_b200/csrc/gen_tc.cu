// bf16 tensor-core path of the MM-GAN generators (Generator / BeatGenerator, network_tests.py:58-123):
// every [Linear -> BatchNorm1d -> Sigmoid] block (network_tests.py:75-80) as ONE tcgen05 GEMM kernel per layer with the
// normalisation folded into its two ends:
//
//   prologue  the A tile (128 batch rows x K input features, bf16, 128-byte-swizzled K-major) is BUILT IN SHARED MEMORY by
//             the epilogue threads from the previous layer's fp32 pre-activations z: a = sigmoid(z * scale + shift), with
//             scale / shift derived per CTA from that layer's batch sums (training) or running stats (eval).  The first
//             layer builds it from the two raw inputs instead (the torch.cat of network_tests.py:86,122 never exists).
//   GEMM      tcgen05.mma (M = 128, N <= 256, K = 16 per instruction), weights bf16 [N][Kp] streamed by TMA, fp32
//             accumulators double-buffered in TMEM so that the MMAs of item i+1 run under the epilogue of item i.
//   epilogue  + bias, then any of: store z (fp32), accumulate the per-column batch sums (sum z, sum z^2: in-tile
//             shuffle transpose-reduction in fp32, cross-tile atomics in fp64), or y = sigmoid(BN(z)) stored fp32.
//
// Training-mode BatchNorm needs the statistics of the whole batch before anything can be normalised, so a layer's z is
// written (hidden layers: <= 1 KB per sample) and normalised by the NEXT layer's prologue; the last layer (4096 wide for
// the Generator = 16 KB per sample) is never written un-normalised: it runs twice, first accumulating only the sums, then
// recomputing the GEMM (K = 64) and writing y once.  Running statistics follow torch (momentum, unbiased variance).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/mmgan_b200.h"

namespace {

// warp 0 TMA, warp 1 MMA, then WG groups of four builder / epilogue warps (one warp per TMEM lane quadrant in each group; the groups take
// alternate 32-column chunks of an accumulator and alternate feature slices of the A tile).  WG = 2 or 4: the epilogue is latency-bound
// (MUFU, tcgen05.ld, the staging barriers), so more resident warps is what speeds it up.
constexpr int GT_MAX_K = 256;
constexpr int GT_A_CHUNK = 128 * 128;           // 128 rows x 64 bf16
constexpr int GT_W_STAGES = 2;
constexpr int GT_Y_STAGE = 128 * 128;           // 128 rows x 32 fp32

struct GenDev {
    const float* x0; const float* x1; int k0, k1;
    int in_mode;                                // 0 raw, 1 BN(batch sums)+sigmoid, 2 BN(running)+sigmoid
    const double* in_sums; const float* in_gamma; const float* in_beta; float* in_run_mean; float* in_run_var;
    const float* bias; int N, NG, n_groups, K, Kp;
    float* z_out; double* out_sums; float* y_out;
    int out_mode;                               // for y_out: 1 batch sums (y_sums), 2 running stats
    const double* y_sums; const float* out_gamma; const float* out_beta; float* out_run_mean; float* out_run_var;
    float momentum, eps; int update_running;
    long long M; int row_tiles;
    int w_stages;                               // weight ring depth: 1 when every CTA has a single item (the hidden layers at B <= 148 row tiles), else 2
    int tm_cols, acc_stride;                    // TMEM columns allocated (one accumulator of NG columns, or two) and the distance between the two
    int col_cache;                              // y_out only: scale / shift of every column group of this CTA are derived ONCE, before its first item
    double cnt;                                 // rows behind the batch sums (= M, or the global batch under SyncBN)
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
// sigmoid(x * scale + shift) with the two MUFU approximations (ex2, rcp: ~2 ulp), branch-free so that the 32-64 independent
// evaluations of a row interleave; nl2e_* are scale / shift pre-multiplied by -log2(e)
__device__ __forceinline__ float fast_sigmoid_affine(float x, float nl2e_scale, float nl2e_shift) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(x, nl2e_scale, nl2e_shift)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return r;
}
constexpr float NEG_LOG2E = -1.4426950408889634f;

// lane l ends up with sum over the warp's lanes of v[l]  (31 shuffles; v is clobbered)
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

__device__ __forceinline__ void bn_scale_shift(double s1, double s2, double count, float gamma, float beta, float eps, float& scale, float& shift,
                                               float& mean_f, float& var_f) {
    const double mu = s1 / count;
    double var = s2 / count - mu * mu;
    if (var < 0) var = 0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    scale = gamma * invstd;
    shift = beta - (float)mu * scale;
    mean_f = (float)mu;
    var_f = (float)var;
}

#ifdef MMG_ABLATION
__device__ long long g_gen_stamps[3][32][4];      // [role: producer / MMA / worker 0][item of CTA 0][event]
#define GEN_STAMP(role, it, ev) do { if (blockIdx.x == 0 && (it) < 32) g_gen_stamps[role][it][ev] = clock64(); } while (0)
#else
#define GEN_STAMP(role, it, ev) do {} while (0)
#endif

template <int WG>
__global__ void __launch_bounds__(64 + WG * 128, 1) gen_layer_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_y,
                                                                        const GenDev a) {
    constexpr int GT_WORKERS = WG * 128, GT_THREADS = 64 + GT_WORKERS;
    constexpr int PER = 64 / WG;                // features of a 64-feature chunk built by one thread
    constexpr int NSTG = 4 / WG;                // output staging buffers per warp group (4 x 16 KB per CTA either way)
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t wfull[GT_W_STAGES], wempty[GT_W_STAGES], aready, tfull[2], tempty[2];
    __shared__ uint32_t tmem_s;
    __shared__ float in_scale[GT_MAX_K], in_shift[GT_MAX_K];
    __shared__ __align__(16) float ep_bias[256], ep_scale[256], ep_shift[256];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int kchunks = a.Kp / 64;
    unsigned char* smem_a = smem;                                   // kchunks x 16 KB
    const int w_stage_bytes = kchunks * a.NG * 128;
    unsigned char* smem_w = smem + kchunks * GT_A_CHUNK;            // w_stages x w_stage_bytes
    unsigned char* smem_y = smem_w + a.w_stages * w_stage_bytes;   // WG warp groups x NSTG x 16 KB output staging (only when y_out)
    float* smem_part = reinterpret_cast<float*>(smem_y + (a.y_out ? 4 * GT_Y_STAGE : 0));      // [4 lane quadrants][sum | sumsq][NG] (only when out_sums)
    float* col_scale = smem_part;                                   // col_cache: [n_groups * NG] scale then [n_groups * NG] shift (never together with out_sums)
    float* col_shift = smem_part + a.n_groups * a.NG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // items = (row tile, column group) pairs, column group fastest; a CTA owns a CONTIGUOUS range, so it changes its row tile (= rebuilds
    // its A operand) at most once or twice per launch
    const int items = a.row_tiles * a.n_groups;
    const int item_begin = (int)((long long)blockIdx.x * items / gridDim.x), item_end = (int)((long long)(blockIdx.x + 1) * items / gridDim.x);

    if (threadIdx.x == 0) {
        for (int i = 0; i < GT_W_STAGES; ++i) { tc::mbar_init(&wfull[i], 1); tc::mbar_init(&wempty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], GT_WORKERS / 32); }
        tc::mbar_init(&aready, GT_WORKERS);
        tc::fence_barrier_init();
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_s, (uint32_t)a.tm_cols); tc::tmem_relinquish(); }
    // input-side BatchNorm folded to scale/shift (every CTA needs all K features; CTA 0 also updates the running stats)
    if (a.in_mode != 0) {
        for (int k = threadIdx.x; k < a.K; k += GT_THREADS) {
            float sc, sh;
            if (a.in_mode == 1) {
                float mu, var;
                bn_scale_shift(a.in_sums[k], a.in_sums[a.K + k], a.cnt, a.in_gamma[k], a.in_beta[k], a.eps, sc, sh, mu, var);
                if (a.update_running && blockIdx.x == 0 && a.in_run_mean) {
                    const float unb = var * (float)(a.cnt / (a.cnt - 1.0));
                    a.in_run_mean[k] = (1.f - a.momentum) * a.in_run_mean[k] + a.momentum * mu;
                    a.in_run_var[k] = (1.f - a.momentum) * a.in_run_var[k] + a.momentum * unb;
                }
            } else {
                const float invstd = rsqrtf(a.in_run_var[k] + a.eps);
                sc = a.in_gamma[k] * invstd;
                sh = a.in_beta[k] - a.in_run_mean[k] * sc;
            }
            in_scale[k] = sc * NEG_LOG2E;                          // pre-multiplied for fast_sigmoid_affine
            in_shift[k] = sh * NEG_LOG2E;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&map_w);
            int it = 0;
            for (int item = item_begin; item < item_end; ++item, ++it) {
                const int stage = it % a.w_stages, phase = (it / a.w_stages) & 1;
                const int grp = item % a.n_groups;
                GEN_STAMP(0, it, 0);
                tc::mbar_wait(&wempty[stage], phase ^ 1);
                GEN_STAMP(0, it, 1);
                tc::mbar_expect_tx(&wfull[stage], w_stage_bytes);
                for (int c = 0; c < kchunks; ++c)
                    tc::tma_load_2d(smem_w + stage * w_stage_bytes + c * a.NG * 128, &map_w, &wfull[stage], c * 64, grp * a.NG);
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint64_t KM128 = tc::smem_desc_base(0, 1024, tc::SW_128B);
            const uint32_t idesc = tc::idesc_bf16(128, (uint32_t)a.NG);
            const uint32_t a_addr = tc::smem_u32(smem_a), w_addr = tc::smem_u32(smem_w);
            int it = 0;
            for (int item = item_begin; item < item_end; ++item, ++it) {
                const int stage = it % a.w_stages, phase = (it / a.w_stages) & 1;
                const int acc = it & 1, acc_phase = (it >> 1) & 1;
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                GEN_STAMP(1, it, 0);
                tc::mbar_wait(&aready, it & 1);
                GEN_STAMP(1, it, 1);
                tc::mbar_wait(&wfull[stage], phase);
                GEN_STAMP(1, it, 2);
                tc::tc_fence_after();
                const uint32_t wb = w_addr + stage * w_stage_bytes;
                for (int c = 0; c < kchunks; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::mma_f16_ss(tmem + acc * a.acc_stride, tc::smem_desc(KM128, a_addr + c * GT_A_CHUNK + k * 32),
                                       tc::smem_desc(KM128, wb + c * a.NG * 128 + k * 32), idesc, (c | k) != 0);
                tc::mma_commit(&wempty[stage]);
                tc::mma_commit(&tfull[acc]);
                GEN_STAMP(1, it, 3);
            }
        }
    } else {
        const int q = warp & 3, h = (warp - 2) >> 2, t = q * 32 + lane;     // t = row of the tile = TMEM lane; h = which half of the columns / features
        const int et = threadIdx.x - 64;                                    // 0 .. GT_WORKERS - 1 over the builder/epilogue threads
        const int eg = et & 127;                                            // index inside the warp group h
        // ---- A-tile builder: row t of row tile `rt`, features [PER h, PER h + PER) of every 64-feature chunk, swizzled 16-byte pieces
        auto build = [&](int rt) {
            const long long row = (long long)rt * 128 + t;
            const bool live = row < a.M;
            for (int c = 0; c < kchunks; ++c) {
                const uint32_t dst = tc::smem_u32(smem_a) + c * GT_A_CHUNK + t * 128;
                const int kb = c * 64 + PER * h;
                float v[PER];                                      // all of the loads are issued before any is used
                if (a.in_mode == 0) {
#pragma unroll
                    for (int e = 0; e < PER; ++e) {
                        const int kk = kb + e;
                        v[e] = !live ? 0.f : (kk < a.k0 ? a.x0[row * a.k0 + kk] : (kk < a.K ? a.x1[row * a.k1 + (kk - a.k0)] : 0.f));
                    }
                } else {                                           // K is a multiple of 8 here (checked on the host)
#pragma unroll
                    for (int j = 0; j < PER / 4; ++j) {
                        const int k = kb + j * 4;
                        const float4 p = (live && k < a.K) ? *reinterpret_cast<const float4*>(a.x0 + row * a.K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[4 * j] = p.x; v[4 * j + 1] = p.y; v[4 * j + 2] = p.z; v[4 * j + 3] = p.w;
                    }
#pragma unroll
                    for (int e = 0; e < PER; ++e) {
                        const int k = kb + e;
                        v[e] = (live && k < a.K) ? fast_sigmoid_affine(v[e], in_scale[k], in_shift[k]) : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < PER / 8; ++j)
                    tc::sts128(dst + ((((PER / 8) * h + j) ^ (t & 7)) << 4), make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                                                                pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7])));
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&aready);
        };
        // output-side BatchNorm of column n folded to y = sigmoid(acc * scale + shift) (bias folded in; both pre-multiplied by -log2 e);
        // `owner` = this call makes the one running-statistics update of the column
        auto column_constants = [&](int n, bool owner, float bias, float& scale, float& shift) {
            float sc = 0.f, sh = 0.f;
            if (n < a.N) {
                if (a.out_mode == 1) {
                    float mu, var;
                    bn_scale_shift(a.y_sums[n], a.y_sums[a.N + n], a.cnt, a.out_gamma[n], a.out_beta[n], a.eps, sc, sh, mu, var);
                    if (a.update_running && owner && a.out_run_mean) {
                        const float unb = var * (float)(a.cnt / (a.cnt - 1.0));
                        a.out_run_mean[n] = (1.f - a.momentum) * a.out_run_mean[n] + a.momentum * mu;
                        a.out_run_var[n] = (1.f - a.momentum) * a.out_run_var[n] + a.momentum * unb;
                    }
                } else {
                    const float invstd = rsqrtf(a.out_run_var[n] + a.eps);
                    sc = a.out_gamma[n] * invstd;
                    sh = a.out_beta[n] - a.out_run_mean[n] * sc;
                }
            }
            scale = sc * NEG_LOG2E;                                  // (acc + b) * sc + sh = acc * sc + (b * sc + sh)
            shift = (bias * sc + sh) * NEG_LOG2E;
        };
        int it = 0, ychunk = 0;
        int item = item_begin;
        if (item < item_end) build(item / a.n_groups);
        if (a.col_cache) {
            // the per-item derivation (five dependent global loads + fp64 arithmetic between two CTA-wide barriers) cost 4.4 K of an item's 15 K
            // cycles.  The groups of a contiguous item range are those of its first min(items, n_groups) items, and exactly one CTA meets
            // (row tile 0, group g) = the owner of the running-statistics update.  Four columns per thread and step: all loads of a step are
            // issued before the first fp64 operation; reciprocal count and rsqrt instead of two fp64 divisions and a square root per column.
            // (A variant that derived the next item's group inside every item's epilogue needed a CTA-wide barrier per item, which couples the
            // warp groups again: 96 us instead of 81.)
            const int total = min(item_end - item_begin, a.n_groups) * a.NG;
            const double inv_cnt = 1.0 / a.cnt;
            const float unb_f = (float)(a.cnt / (a.cnt - 1.0));
            for (int base = et; base < total; base += 4 * GT_WORKERS) {
                double p0[4], p1[4];
                float ga[4], be[4], bi[4], rm[4], rv[4];
                int nn[4];
                bool own[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * GT_WORKERS;
                    const int itm = item_begin + i / a.NG, n = (itm % a.n_groups) * a.NG + i % a.NG;
                    nn[u] = (i < total) ? n : -1;
                    own[u] = a.out_mode == 1 && a.update_running && a.out_run_mean && itm / a.n_groups == 0;
                    const bool live = i < total && n < a.N;
                    p0[u] = p1[u] = 0.0; ga[u] = be[u] = bi[u] = rm[u] = rv[u] = 0.f;
                    if (live) {
                        if (a.out_mode == 1) { p0[u] = a.y_sums[n]; p1[u] = a.y_sums[a.N + n]; }
                        if (a.out_mode != 1 || own[u]) { rm[u] = a.out_run_mean[n]; rv[u] = a.out_run_var[n]; }
                        ga[u] = a.out_gamma[n]; be[u] = a.out_beta[n];
                        bi[u] = a.bias ? a.bias[n] : 0.f;
                    } else own[u] = false;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (nn[u] < 0) continue;
                    float sc = 0.f, sh = 0.f;
                    if (nn[u] < a.N) {
                        if (a.out_mode == 1) {
                            const double mu = p0[u] * inv_cnt;
                            double var = p1[u] * inv_cnt - mu * mu;
                            if (var < 0) var = 0;
                            sc = ga[u] * (float)rsqrt(var + (double)a.eps);
                            sh = be[u] - (float)mu * sc;
                            if (own[u]) {
                                a.out_run_mean[nn[u]] = (1.f - a.momentum) * rm[u] + a.momentum * (float)mu;
                                a.out_run_var[nn[u]] = (1.f - a.momentum) * rv[u] + a.momentum * ((float)var * unb_f);
                            }
                        } else {
                            sc = ga[u] * rsqrtf(rv[u] + a.eps);
                            sh = be[u] - rm[u] * sc;
                        }
                    }
                    col_scale[nn[u]] = sc * NEG_LOG2E;                // (acc + b) * sc + sh = acc * sc + (b * sc + sh)
                    col_shift[nn[u]] = (bi[u] * sc + sh) * NEG_LOG2E;
                }
            }
            asm volatile("bar.sync 7, %0;" ::"n"(GT_WORKERS) : "memory");
        }
        for (; item < item_end; ++item, ++it) {
            const int acc = it & 1, acc_phase = (it >> 1) & 1;
            const int rt = item / a.n_groups, grp = item % a.n_groups, n0 = grp * a.NG;
            // per-item column constants (bias; output-side BN folded to scale/shift) unless they were derived up front
            const float* scp = ep_scale;
            const float* shp = ep_shift;
            if (a.col_cache) { scp = col_scale + n0; shp = col_shift + n0; }
            else {
                asm volatile("bar.sync 7, %0;" ::"n"(GT_WORKERS) : "memory");      // previous item's epilogue no longer reads ep_*
                for (int c = et; c < a.NG; c += GT_WORKERS) {
                    const int n = n0 + c;
                    ep_bias[c] = (n < a.N && a.bias) ? a.bias[n] : 0.f;
                    if (a.y_out) column_constants(n, rt == 0, ep_bias[c], ep_scale[c], ep_shift[c]);      // exactly one item owns (row tile 0, column n)
                }
                asm volatile("bar.sync 7, %0;" ::"n"(GT_WORKERS) : "memory");
            }
            if (et == 0) GEN_STAMP(2, it, 0);
            tc::mbar_wait(&tfull[acc], acc_phase);
            if (et == 0) GEN_STAMP(2, it, 1);
            tc::tc_fence_after();
            // the MMAs of this item are complete: the A buffer is free, build the next item's tile so its MMAs overlap this epilogue
            if (item + 1 < item_end) {
                if ((item + 1) / a.n_groups != rt) build((item + 1) / a.n_groups);
                else tc::mbar_arrive(&aready);                       // same row tile: the operand in shared memory stays
            }
            const long long row = (long long)rt * 128 + t;
            const bool live = row < a.M;
            for (int c0 = 32 * h; c0 < a.NG; c0 += 32 * WG) {         // the WG warps of a lane quadrant take alternate 32-column chunks
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + acc * a.acc_stride + c0, r);
                tc::tmem_ld_wait();
                float z[32];
                if (a.z_out || a.out_sums) {
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const float4 bb = reinterpret_cast<const float4*>(ep_bias + c0)[e4];
                        z[4 * e4] = __uint_as_float(r[4 * e4]) + bb.x; z[4 * e4 + 1] = __uint_as_float(r[4 * e4 + 1]) + bb.y;
                        z[4 * e4 + 2] = __uint_as_float(r[4 * e4 + 2]) + bb.z; z[4 * e4 + 3] = __uint_as_float(r[4 * e4 + 3]) + bb.w;
                    }
                }
                const int ncols = a.N - (n0 + c0);                  // columns of this chunk that exist (>= 32: all)
                if (a.z_out && live) {
                    float* zp = a.z_out + row * a.N + n0 + c0;
                    if (ncols >= 32 && (a.N & 3) == 0) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(zp + e) = make_float4(z[e], z[e + 1], z[e + 2], z[e + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) if (e < ncols) zp[e] = z[e];
                    }
                }
                if (a.y_out) {
                    // y tile (128 rows x 32 fp32) -> 128-byte-swizzled staging buffer -> one TMA store (rows >= M and columns >= N are clipped)
                    unsigned char* stg = smem_y + (NSTG * h + (ychunk % NSTG)) * GT_Y_STAGE;
                    if (eg == 0) tc::bulk_wait_group_read<NSTG - 1>(); // the store this warp group issued from this buffer NSTG chunks ago has read it
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
                    float y[32];
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const float4 sc = reinterpret_cast<const float4*>(scp + c0)[e4], sh = reinterpret_cast<const float4*>(shp + c0)[e4];
                        y[4 * e4] = fast_sigmoid_affine(__uint_as_float(r[4 * e4]), sc.x, sh.x);
                        y[4 * e4 + 1] = fast_sigmoid_affine(__uint_as_float(r[4 * e4 + 1]), sc.y, sh.y);
                        y[4 * e4 + 2] = fast_sigmoid_affine(__uint_as_float(r[4 * e4 + 2]), sc.z, sh.z);
                        y[4 * e4 + 3] = fast_sigmoid_affine(__uint_as_float(r[4 * e4 + 3]), sc.w, sh.w);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        tc::sts128(tc::smem_u32(stg) + t * 128 + ((j ^ (t & 7)) << 4), make_uint4(__float_as_uint(y[4 * j]), __float_as_uint(y[4 * j + 1]),
                                                                                               __float_as_uint(y[4 * j + 2]), __float_as_uint(y[4 * j + 3])));
                    tc::fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
                    if (eg == 0) {
                        tc::tma_store_2d(&map_y, stg, n0 + c0, rt * 128);
                        tc::bulk_commit_group();
                    }
                    ++ychunk;
                }
                if (a.out_sums) {
                    float s1[32], s2[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) { s1[e] = live ? z[e] : 0.f; s2[e] = s1[e] * s1[e]; }
                    const float c1 = colsum32(s1, lane), c2 = colsum32(s2, lane);
                    smem_part[(2 * q) * a.NG + c0 + lane] = c1;      // 32-row partial of this warp; the four quadrants are combined below
                    smem_part[(2 * q + 1) * a.NG + c0 + lane] = c2;
                }
            }
            if (a.out_sums) {
                // one fp64 atomic per column and CTA instead of one per warp: same-address atomics serialise in L2 (128 row tiles x 4 warps on
                // each of the 2 N addresses cost ~15 us per layer; the in-CTA combine of the four 32-row partials is done in fp64 as before)
                asm volatile("bar.sync 7, %0;" ::"n"(GT_WORKERS) : "memory");
                for (int c = et; c < 2 * a.NG; c += GT_WORKERS) {
                    const int which = c >= a.NG ? 1 : 0, col = c - which * a.NG;
                    if (n0 + col < a.N) {
                        const double v = ((double)smem_part[which * a.NG + col] + (double)smem_part[(2 + which) * a.NG + col]) +
                                         ((double)smem_part[(4 + which) * a.NG + col] + (double)smem_part[(6 + which) * a.NG + col]);
                        atomicAdd(&a.out_sums[which * a.N + n0 + col], v);
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (et == 0) GEN_STAMP(2, it, 2);
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);

        }
        if (eg == 0) tc::bulk_wait_group_read<0>();                // staging buffers must outlive the last TMA store's reads
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, (uint32_t)a.tm_cols);
}

__global__ void gen_pack_weights_kernel(const float* __restrict__ w, int N, int K, int Np, int Kp, __nv_bfloat16* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np * Kp) return;
    const int n = i / Kp, k = i - n * Kp;
    out[i] = __float2bfloat16((n < N && k < K) ? w[(size_t)n * K + k] : 0.f);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int g_worker_groups = 4;                           // builder / epilogue warp groups of gen_layer_tc_kernel (mmg_gen_set_worker_groups)

// ------------------------------------------------------------------------------------------------
// Analytic batch statistics of a WIDE layer fed by a NARROW one (the generator's 64 -> 4096 output layer): z = a W^T + b, so
//   sum_r z[r][n]   = w_n . s + M b_n                       with s = sum_r a_r            (K values)
//   sum_r z[r][n]^2 = w_n^T G w_n + 2 b_n (w_n . s) + M b_n^2   with G = sum_r a_r a_r^T  (K x K Gram matrix)
// i.e. 64 x 64 + 64 numbers about the input replace a whole GEMM pass whose only product was the column sums (M x 4096 accumulators read back
// from TMEM and squared).  a and w are the SAME bf16-rounded operands the tensor-core pass multiplies; sums are accumulated in fp64.
// ------------------------------------------------------------------------------------------------
constexpr int GS_K = 64;                           // input features handled by the Gram path
constexpr int GS_ROWS = 64;                        // rows per shared-memory tile
constexpr int GS_PART = GS_K * GS_K + GS_K;        // doubles per partial result: G then s

__global__ void __launch_bounds__(256) gen_gram_partial_kernel(const float* __restrict__ z, long long M, int K, const double* __restrict__ in_sums,
                                                               double count, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                               double* __restrict__ part) {
    __shared__ float at[GS_ROWS][GS_K + 1];
    __shared__ float sc[GS_K], sh[GS_K];
    const int tid = threadIdx.x;
    if (tid < GS_K) {
        float s = 0.f, h = 0.f, mu, var;
        if (tid < K) bn_scale_shift(in_sums[tid], in_sums[K + tid], count, gamma[tid], beta[tid], eps, s, h, mu, var);
        sc[tid] = s * NEG_LOG2E; sh[tid] = h * NEG_LOG2E;
    }
    __syncthreads();
    const int i0 = (tid >> 4) * 4, j0 = (tid & 15) * 4;              // this thread's 4 x 4 block of G
    double acc[16], ssum = 0.0;
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.0;
    const long long tiles = (M + GS_ROWS - 1) / GS_ROWS;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        {   // a = bf16(sigmoid(BN(z))) exactly as the tensor-core kernel builds its A operand; 16 values per thread
            const int r = tid >> 2, k0 = (tid & 3) * 16;
            const long long row = tile * GS_ROWS + r;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int k = k0 + e;
                float v = 0.f;
                if (row < M && k < K) v = __bfloat162float(__float2bfloat16(fast_sigmoid_affine(z[row * K + k], sc[k], sh[k])));
                at[r][k] = v;
            }
        }
        __syncthreads();
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = 0.f;
#pragma unroll 8
        for (int r = 0; r < GS_ROWS; ++r) {
            float ai[4], aj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { ai[u] = at[r][i0 + u]; aj[u] = at[r][j0 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) f[4 * u + v] = fmaf(ai[u], aj[v], f[4 * u + v]);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] += (double)f[e];        // fp32 inside a 64-row tile (values in (0,1)), fp64 across tiles
        if (tid < GS_K) {
            float c = 0.f;
            for (int r = 0; r < GS_ROWS; ++r) c += at[r][tid];
            ssum += (double)c;
        }
        __syncthreads();
    }
    double* out = part + (size_t)blockIdx.x * GS_PART;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) out[(i0 + u) * GS_K + j0 + v] = acc[4 * u + v];
    if (tid < GS_K) out[GS_K * GS_K + tid] = ssum;
}

// 8 threads per output value (partials p = part, part + 8, ...), shuffle reduction: 4160 outputs x 8 threads = 130 CTAs of 256 threads
// (the round-1 version used one thread per output in 17 CTAs: 27 us of load latency for 4.9 MB)
__global__ void __launch_bounds__(256) gen_gram_reduce_kernel(const double* __restrict__ part, int nparts, double* __restrict__ red) {
    const int i = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
    const int ii = i < GS_PART ? i : GS_PART - 1;                  // keep the whole warp in the shuffles
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int p = sub;
    for (; p + 24 < nparts; p += 32) {                             // 4 independent loads in flight per thread
        a0 += part[(size_t)p * GS_PART + ii];
        a1 += part[(size_t)(p + 8) * GS_PART + ii];
        a2 += part[(size_t)(p + 16) * GS_PART + ii];
        a3 += part[(size_t)(p + 24) * GS_PART + ii];
    }
    for (; p < nparts; p += 8) a0 += part[(size_t)p * GS_PART + ii];
    double v = (a0 + a1) + (a2 + a3);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (sub == 0 && i < GS_PART) red[i] = v;
}

// four threads per output column: thread `part` takes the rows i = part, part + 4, ... of G; quad shuffle reduction
__global__ void __launch_bounds__(256) gen_gram_colstats_kernel(const double* __restrict__ red, const float* __restrict__ w, const float* __restrict__ bias,
                                                                int N, int K, double count, double* __restrict__ out_sums) {
    constexpr int GP = GS_K + 1;                                     // padded row pitch: the four threads of a column read four different banks
    __shared__ double Gs[GS_K * GP], ss[GS_K];
    for (int i = threadIdx.x; i < GS_K * GS_K; i += blockDim.x) Gs[(i / GS_K) * GP + (i % GS_K)] = red[i];
    if (threadIdx.x < GS_K) ss[threadIdx.x] = red[GS_K * GS_K + threadIdx.x];
    __syncthreads();
    const int n = blockIdx.x * 64 + (threadIdx.x >> 2), part = threadIdx.x & 3;
    const int nn = n < N ? n : N - 1;                                // keep the whole warp in the shuffles
    float wn[GS_K];
#pragma unroll
    for (int k = 0; k < GS_K; ++k) wn[k] = k < K ? __bfloat162float(__float2bfloat16(w[(size_t)nn * K + k])) : 0.f;      // the bf16 operand of the GEMM
    double q = 0.0, l = 0.0;
#pragma unroll
    for (int ii = 0; ii < GS_K / 4; ++ii) {
        const int i = 4 * ii + part;
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < GS_K; ++j) t = fma(Gs[i * GP + j], (double)wn[j], t);
        const double wi = (double)(part == 0 ? wn[4 * ii] : part == 1 ? wn[4 * ii + 1] : part == 2 ? wn[4 * ii + 2] : wn[4 * ii + 3]);
        q = fma(t, wi, q);
        l = fma(ss[i], wi, l);
    }
    q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2);
    l += __shfl_xor_sync(0xffffffffu, l, 1); l += __shfl_xor_sync(0xffffffffu, l, 2);
    if (n < N && part == 0) {
        const double b = bias ? (double)bias[n] : 0.0;
        out_sums[n] = l + count * b;
        out_sums[N + n] = q + 2.0 * b * l + count * b * b;
    }
}

// ------------------------------------------------------------------------------------------------
// The three hidden [Linear -> BatchNorm1d -> Sigmoid] blocks of a generator (network_tests.py:68-71,103-106) in ONE cooperative launch, train mode:
// one CTA per 128-row tile keeps its rows' pre-activations in TMEM (layer 1: columns 0.., layer 2 and 3 behind it) from layer to layer; a
// layer's column sums leave the CTA as one fp64 atomic per column, a grid-wide barrier makes the batch statistics complete, and the next
// layer's A operand is built from the TMEM accumulators exactly as gen_layer_tc_kernel builds it from the stored z (z = acc + bias in fp32,
// then sigmoid(z * scale + shift), rounded to bf16).  Only the last hidden layer's z is written (the output layer's kernel reads it), and
// the CTA's part of the output layer's Gram statistics (G = sum a a^T, s = sum a over its rows) is formed from the same TMEM columns.
// All three weight matrices are fetched by TMA at kernel start.  3 launches + the Gram partial kernel become 1.
struct HiddenDev {
    const float* x0; const float* x1; int k0, k1;
    const float* bias[3]; const float* gamma[3]; const float* beta[3]; float* run_mean[3]; float* run_var[3]; double* sums[3];
    int N[3], NG[3], K[3], Kp[3], tcol[3];
    int w_off[3];                               // byte offsets of the weight tiles behind the A buffer
    int a_bytes, part_off;
    float* z_out; double* gram_part; unsigned int* barrier;
    float momentum, eps; int update_running; long long M; double cnt;
};

constexpr int GH_WORKERS = 256, GH_THREADS = 64 + GH_WORKERS;

__global__ void __launch_bounds__(GH_THREADS, 1) gen_hidden_fused_kernel(const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1,
                                                                         const __grid_constant__ CUtensorMap map_w2, const HiddenDev a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t wbar[3], aready, accfull;
    __shared__ uint32_t tmem_s;
    __shared__ float in_scale[GT_MAX_K], in_shift[GT_MAX_K], in_bias[GT_MAX_K];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* smem_a = smem;
    float* smem_part = reinterpret_cast<float*>(smem + a.part_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rt = blockIdx.x;

    if (threadIdx.x == 0) {
        for (int l = 0; l < 3; ++l) tc::mbar_init(&wbar[l], 1);
        tc::mbar_init(&aready, GH_WORKERS);
        tc::mbar_init(&accfull, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_s, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp == 0) {
        if (tc::elect_one()) {
            const CUtensorMap* maps[3] = {&map_w0, &map_w1, &map_w2};
            for (int l = 0; l < 3; ++l) {
                tc::tma_prefetch_desc(maps[l]);
                const int kch = a.Kp[l] / 64;
                tc::mbar_expect_tx(&wbar[l], (uint32_t)(kch * a.NG[l] * 128));
                for (int c = 0; c < kch; ++c) tc::tma_load_2d(smem + a.w_off[l] + c * a.NG[l] * 128, maps[l], &wbar[l], c * 64, 0);
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint64_t KM128 = tc::smem_desc_base(0, 1024, tc::SW_128B);
            const uint32_t a_addr = tc::smem_u32(smem_a);
            for (int l = 0; l < 3; ++l) {
                const uint32_t idesc = tc::idesc_bf16(128, (uint32_t)a.NG[l]);
                const uint32_t wb = tc::smem_u32(smem + a.w_off[l]);
                tc::mbar_wait(&aready, (uint32_t)(l & 1));
                tc::mbar_wait(&wbar[l], 0);
                tc::tc_fence_after();
                const int kch = a.Kp[l] / 64;
                for (int c = 0; c < kch; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::mma_f16_ss(tmem + a.tcol[l], tc::smem_desc(KM128, a_addr + c * GT_A_CHUNK + k * 32), tc::smem_desc(KM128, wb + c * a.NG[l] * 128 + k * 32), idesc,
                                       (c | k) != 0);
                tc::mma_commit(&accfull);
            }
        }
    } else {
        const int q = warp & 3, h = (warp - 2) >> 2, t = q * 32 + lane;     // t = row of the tile = TMEM lane; h = which half of a 64-feature chunk / alternate 32-column chunks
        const int et = threadIdx.x - 64;
        const long long row = (long long)rt * 128 + t;
        const bool live = row < a.M;
        auto bar_workers = [&]() { asm volatile("bar.sync 7, %0;" ::"n"(GH_WORKERS) : "memory"); };
        // ---- layer 1 operand from the two raw inputs (the torch.cat of network_tests.py:86,122 never exists)
        {
            const int K = a.K[0], kch = a.Kp[0] / 64;
            for (int c = 0; c < kch; ++c) {
                const uint32_t dst = tc::smem_u32(smem_a) + c * GT_A_CHUNK + t * 128;
                const int kb = c * 64 + 32 * h;
                float v[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int kk = kb + e;
                    v[e] = !live ? 0.f : (kk < a.k0 ? a.x0[row * a.k0 + kk] : (kk < K ? a.x1[row * a.k1 + (kk - a.k0)] : 0.f));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tc::sts128(dst + (((4 * h + j) ^ (t & 7)) << 4), make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                                                                pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7])));
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&aready);
        }
        for (int l = 0; l < 3; ++l) {
            const int N = a.N[l], NG = a.NG[l];
            bar_workers();                                   // the previous layer's operand builders have read in_bias
            for (int c = et; c < NG; c += GH_WORKERS) in_bias[c] = (c < N && a.bias[l]) ? a.bias[l][c] : 0.f;
            bar_workers();
            tc::mbar_wait(&accfull, (uint32_t)(l & 1));
            tc::tc_fence_after();
            // ---- epilogue: z = acc + bias; column sums of z and z^2 over this tile's rows (32-row partials in fp32, combined in fp64); the last layer's z is stored
            for (int c0 = 32 * h; c0 < NG; c0 += 64) {
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + a.tcol[l] + c0, r);
                tc::tmem_ld_wait();
                float s1[32], s2[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float z = __uint_as_float(r[e]) + in_bias[c0 + e];
                    s1[e] = live ? z : 0.f;
                    s2[e] = s1[e] * s1[e];
                }
                if (l == 2 && a.z_out && live) {
                    float* zp = a.z_out + row * N + c0;
                    if (c0 + 32 <= N && (N & 3) == 0) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(zp + e) = make_float4(s1[e], s1[e + 1], s1[e + 2], s1[e + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) if (c0 + e < N) zp[e] = s1[e];
                    }
                }
                const float c1 = colsum32(s1, lane), c2 = colsum32(s2, lane);
                smem_part[(2 * q) * NG + c0 + lane] = c1;
                smem_part[(2 * q + 1) * NG + c0 + lane] = c2;
            }
            bar_workers();
            for (int c = et; c < 2 * NG; c += GH_WORKERS) {
                const int which = c >= NG ? 1 : 0, col = c - which * NG;
                if (col < N) {
                    const double v = ((double)smem_part[which * NG + col] + (double)smem_part[(2 + which) * NG + col]) +
                                     ((double)smem_part[(4 + which) * NG + col] + (double)smem_part[(6 + which) * NG + col]);
                    atomicAdd(&a.sums[l][which * N + col], v);
                }
            }
            // ---- grid-wide barrier: every CTA's atomics are performed before any CTA reads the sums (cooperative launch: all CTAs are resident)
            bar_workers();
            if (et == 0) {
                __threadfence();
                atomicAdd(a.barrier, 1u);
                const unsigned int target = gridDim.x * (unsigned int)(l + 1);
                while (*reinterpret_cast<volatile unsigned int*>(a.barrier) < target) {}
                __threadfence();
            }
            bar_workers();
            // ---- this layer's BatchNorm folded to scale / shift (every CTA needs all N features; CTA 0 updates the running statistics)
            for (int k = et; k < N; k += GH_WORKERS) {
                float sc, sh, mu, var;
                bn_scale_shift(__ldcg(&a.sums[l][k]), __ldcg(&a.sums[l][N + k]), a.cnt, a.gamma[l][k], a.beta[l][k], a.eps, sc, sh, mu, var);
                if (a.update_running && blockIdx.x == 0 && a.run_mean[l]) {
                    const float unb = var * (float)(a.cnt / (a.cnt - 1.0));
                    a.run_mean[l][k] = (1.f - a.momentum) * a.run_mean[l][k] + a.momentum * mu;
                    a.run_var[l][k] = (1.f - a.momentum) * a.run_var[l][k] + a.momentum * unb;
                }
                in_scale[k] = sc * NEG_LOG2E;
                in_shift[k] = sh * NEG_LOG2E;
            }
            bar_workers();
            if (l < 2) {
                // ---- next layer's A operand from the TMEM accumulators: features [64 c + 32 h, + 32) of row t
                const int kch = a.Kp[l + 1] / 64;                // = ceil(N / 64)
                for (int c = 0; c < kch; ++c) {
                    const int kb = c * 64 + 32 * h;
                    uint32_t r[32];
                    float v[32];
                    if (kb < NG) {
                        tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + a.tcol[l] + kb, r);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = (live && kb + e < N) ? fast_sigmoid_affine(__uint_as_float(r[e]) + in_bias[kb + e], in_scale[kb + e], in_shift[kb + e]) : 0.f;
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = 0.f;
                    }
                    const uint32_t dst = tc::smem_u32(smem_a) + c * GT_A_CHUNK + t * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        tc::sts128(dst + (((4 * h + j) ^ (t & 7)) << 4), make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                                                                    pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7])));
                }
                tc::tc_fence_before();
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&aready);
            } else if (a.gram_part) {
                // ---- this CTA's part of the output layer's Gram statistics (see gen_gram_partial_kernel): a = bf16(sigmoid(BN(z))) of its 128 rows,
                // two 64-row halves accumulated in fp32, added in fp64.  The tile lives where the layer-1 weights were.
                float (*at)[GS_K + 1] = reinterpret_cast<float (*)[GS_K + 1]>(smem + a.w_off[0]);
                {
                    uint32_t r[32];
                    tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + a.tcol[2] + 32 * h, r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int k = 32 * h + e;
                        at[t][k] = (live && k < N) ? __bfloat162float(__float2bfloat16(fast_sigmoid_affine(__uint_as_float(r[e]) + in_bias[k], in_scale[k], in_shift[k]))) : 0.f;
                    }
                }
                bar_workers();
                const int i0 = (et >> 4) * 4, j0 = (et & 15) * 4;
                double acc[16], ssum = 0.0;
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[e] = 0.0;
                for (int half = 0; half < 2; ++half) {
                    float f[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) f[e] = 0.f;
#pragma unroll 8
                    for (int rr = 0; rr < GS_ROWS; ++rr) {
                        const int r0 = half * GS_ROWS + rr;
                        float ai[4], aj[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) { ai[u] = at[r0][i0 + u]; aj[u] = at[r0][j0 + u]; }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v) f[4 * u + v] = fmaf(ai[u], aj[v], f[4 * u + v]);
                    }
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc[e] += (double)f[e];
                    if (et < GS_K) {
                        float c = 0.f;
                        for (int rr = 0; rr < GS_ROWS; ++rr) c += at[half * GS_ROWS + rr][et];
                        ssum += (double)c;
                    }
                }
                double* out = a.gram_part + (size_t)blockIdx.x * GS_PART;
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) out[(i0 + u) * GS_K + j0 + v] = acc[4 * u + v];
                if (et < GS_K) out[GS_K * GS_K + et] = ssum;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 512);
}


}  // namespace

extern "C" {

// bytes of the bf16 operand copy of an nn.Linear weight (N, K): [Np][Kp], Np = N rounded up to 32 (to 256 when N > 256), Kp = K to 64
size_t mmg_gen_packed_weight_bytes(int N, int K) {
    if (N <= 0 || K <= 0) return 0;
    const int Np = N > 256 ? round_up(N, 256) : round_up(N, 32);
    return (size_t)Np * round_up(K, 64) * 2;
}

int mmg_gen_pack_weight(const float* w, int N, int K, void* packed, void* stream) {
    MMG_REQUIRE(w && packed && N > 0 && K > 0, MMG_EINVAL, "gen_pack_weight: bad arguments");
    const int Np = N > 256 ? round_up(N, 256) : round_up(N, 32), Kp = round_up(K, 64);
    gen_pack_weights_kernel<<<(Np * Kp + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, N, K, Np, Kp, (__nv_bfloat16*)packed);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// workspace of mmg_gen_layer_stats_gram: one partial (G, s) per CTA + the reduced copy
size_t mmg_gen_layer_stats_gram_workspace(void) { return sizeof(double) * (size_t)GS_PART * (MMG_NUM_SMS + 1); }

// Batch sums of z = sigmoid(BN(z_prev)) W^T + b over the M local rows, written to out_sums (sum | sum of squares, fp64 [2][N]) WITHOUT running
// the GEMM: see the Gram-matrix comment above.  z_prev (M,K) fp32 with K <= 64 and K % 8 == 0 is the previous layer's pre-activation, in_sums its
// batch sums over stat_count rows (0 = M); weight (N,K) / bias (N,) are the fp32 master tensors of the layer (rounded to bf16 here exactly as
// mmg_gen_pack_weight does).  Train mode only.
int mmg_gen_layer_stats_gram(const float* z_prev, int64_t M, int K, const double* in_sums, int64_t stat_count, const float* in_gamma,
                             const float* in_beta, float eps, const float* weight, const float* bias, int N, double* out_sums, void* workspace,
                             size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMG_REQUIRE(z_prev && in_sums && in_gamma && in_beta && weight && out_sums && workspace && M > 0 && N > 0, MMG_EINVAL, "gen_layer_stats_gram: bad arguments");
    MMG_REQUIRE(K > 0 && K <= GS_K, MMG_EUNSUPPORTED, "gen_layer_stats_gram: at most %d input features (got %d)", GS_K, K);
    MMG_REQUIRE(ws_bytes >= mmg_gen_layer_stats_gram_workspace(), MMG_EINVAL, "gen_layer_stats_gram: workspace too small");
    const double count = (double)(stat_count > 0 ? stat_count : M);
    MMG_REQUIRE(count > 1, MMG_EINVAL, "Expected more than 1 value per channel when training, got input size (%lld, %d)", (long long)M, K);
    double* part = (double*)workspace;
    double* red = part + (size_t)GS_PART * MMG_NUM_SMS;
    const long long tiles = (M + GS_ROWS - 1) / GS_ROWS;
    const int grid = (int)(tiles < MMG_NUM_SMS ? tiles : MMG_NUM_SMS);
    gen_gram_partial_kernel<<<grid, 256, 0, stream>>>(z_prev, (long long)M, K, in_sums, count, in_gamma, in_beta, eps, part);
    MMG_LAUNCH_CHECK();
    gen_gram_reduce_kernel<<<(GS_PART + 31) / 32, 256, 0, stream>>>(part, grid, red);
    MMG_LAUNCH_CHECK();
    gen_gram_colstats_kernel<<<(N + 63) / 64, 256, 0, stream>>>(red, weight, bias, N, K, (double)M, out_sums);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// ---- the three hidden blocks in one cooperative launch (gen_hidden_fused_kernel); train mode, local batch statistics
static int hidden_layout(const mmg_gen_hidden_args* p, HiddenDev& a, size_t& smem) {
    int k_in = p->k0 + p->k1, tcol = 0, max_kch = 0, max_ng = 0;
    for (int l = 0; l < 3; ++l) {
        if (p->N[l] <= 0 || p->N[l] > 256 || k_in <= 0 || k_in > GT_MAX_K) return 0;
        a.N[l] = p->N[l]; a.NG[l] = round_up(p->N[l], 32); a.K[l] = k_in; a.Kp[l] = round_up(k_in, 64);
        a.tcol[l] = tcol;
        tcol += a.NG[l];
        max_kch = a.Kp[l] / 64 > max_kch ? a.Kp[l] / 64 : max_kch;
        max_ng = a.NG[l] > max_ng ? a.NG[l] : max_ng;
        k_in = p->N[l];
    }
    if (tcol > 512) return 0;
    a.a_bytes = max_kch * GT_A_CHUNK;
    int off = a.a_bytes;
    for (int l = 0; l < 3; ++l) { a.w_off[l] = off; off += (a.Kp[l] / 64) * a.NG[l] * 128; }
    a.part_off = off;
    off += 8 * max_ng * (int)sizeof(float);
    smem = 1024 + (size_t)off;
    if (p->gram_part && ((a.Kp[0] / 64) * a.NG[0] * 128 < (int)(128 * (GS_K + 1) * sizeof(float)) || p->N[2] > GS_K)) return 0;     // the Gram tile reuses the layer-1 weight tile
    return smem <= 220 * 1024;
}

int mmg_gen_hidden_fused_supported(int64_t M, int k_in, const int* N3, int with_gram) {
    if (!N3 || M <= 1 || (M + 127) / 128 > MMG_NUM_SMS) return 0;
    mmg_gen_hidden_args p = {};
    p.k0 = k_in; p.k1 = 0;
    for (int l = 0; l < 3; ++l) p.N[l] = N3[l];
    p.gram_part = with_gram ? (double*)16 : nullptr;
    HiddenDev a;
    size_t smem;
    return hidden_layout(&p, a, smem);
}

int mmg_gen_hidden_fused(const mmg_gen_hidden_args* p, void* stream) {
    MMG_REQUIRE(p && p->x0 && p->M > 1 && p->k0 > 0 && p->k1 >= 0 && (p->k1 == 0 || p->x1) && p->barrier && p->z_out, MMG_EINVAL, "gen_hidden_fused: bad arguments");
    MMG_REQUIRE(p->stat_count == 0 || p->stat_count == p->M, MMG_EUNSUPPORTED, "gen_hidden_fused: local batch statistics only (SyncBN runs the per-layer kernels)");
    for (int l = 0; l < 3; ++l)
        MMG_REQUIRE(p->w_packed[l] && p->gamma[l] && p->beta[l] && p->sums[l], MMG_EINVAL, "gen_hidden_fused: missing tensors of layer %d", l);
    HiddenDev a;
    size_t smem;
    const long long row_tiles = (p->M + 127) / 128;
    MMG_REQUIRE(row_tiles <= MMG_NUM_SMS && hidden_layout(p, a, smem), MMG_EUNSUPPORTED, "gen_hidden_fused: shape does not fit one cooperative launch (see mmg_gen_hidden_fused_supported)");
    a.x0 = p->x0; a.x1 = p->x1; a.k0 = p->k0; a.k1 = p->k1;
    CUtensorMap maps[3];
    for (int l = 0; l < 3; ++l) {
        a.bias[l] = p->bias[l]; a.gamma[l] = p->gamma[l]; a.beta[l] = p->beta[l]; a.run_mean[l] = p->run_mean[l]; a.run_var[l] = p->run_var[l]; a.sums[l] = p->sums[l];
        MMG_REQUIRE(tc::make_map_2d_bf16(&maps[l], p->w_packed[l], (uint64_t)a.Kp[l], (uint64_t)a.NG[l], (uint64_t)a.Kp[l] * 2, 64, (uint32_t)a.NG[l], CU_TENSOR_MAP_SWIZZLE_128B) == 0,
                    MMG_EINVAL, "gen_hidden_fused: cuTensorMapEncodeTiled(w%d) failed", l);
    }
    a.z_out = p->z_out; a.gram_part = p->gram_part; a.barrier = p->barrier;
    a.momentum = p->momentum; a.eps = p->eps; a.update_running = p->update_running; a.M = p->M; a.cnt = (double)p->M;
    static bool attr_done = false;
    if (!attr_done) {
        MMG_CUDA(cudaFuncSetAttribute(gen_hidden_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done = true;
    }
    void* args[] = {&maps[0], &maps[1], &maps[2], &a};
    MMG_CUDA(cudaLaunchCooperativeKernel((const void*)gen_hidden_fused_kernel, dim3((unsigned)row_tiles), dim3(GH_THREADS), args, smem, (cudaStream_t)stream));
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// second half of mmg_gen_layer_stats_gram for partials that mmg_gen_hidden_fused wrote (nparts = row tiles of 128): reduce + column statistics
int mmg_gen_layer_stats_gram_finish(int nparts, const float* weight, const float* bias, int N, int K, int64_t M, double* out_sums, void* workspace,
                                    size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMG_REQUIRE(weight && out_sums && workspace && nparts > 0 && nparts <= MMG_NUM_SMS && N > 0 && K > 0 && K <= GS_K && M > 1, MMG_EINVAL, "gen_layer_stats_gram_finish: bad arguments");
    MMG_REQUIRE(ws_bytes >= mmg_gen_layer_stats_gram_workspace(), MMG_EINVAL, "gen_layer_stats_gram_finish: workspace too small");
    double* part = (double*)workspace;
    double* red = part + (size_t)GS_PART * MMG_NUM_SMS;
    gen_gram_reduce_kernel<<<(GS_PART + 31) / 32, 256, 0, stream>>>(part, nparts, red);
    MMG_LAUNCH_CHECK();
    gen_gram_colstats_kernel<<<(N + 63) / 64, 256, 0, stream>>>(red, weight, bias, N, K, (double)M, out_sums);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// process-wide tuning switch: 2 or 4 groups of four builder / epilogue warps per CTA (default 4: G1 forward 151 us against 176 at B = 16 384); returns
// the previous value
int mmg_gen_set_worker_groups(int groups) {
    const int prev = g_worker_groups;
    if (groups == 2 || groups == 4) g_worker_groups = groups;
    return prev;
}

#ifdef MMG_ABLATION
int mmg_gen_get_stamps(long long* host) { return (int)cudaMemcpyFromSymbol(host, g_gen_stamps, sizeof(long long) * 3 * 32 * 4); }
#endif

int mmg_gen_layer_fwd(const mmg_gen_layer_args* p, void* stream) {
    MMG_REQUIRE(p && p->x0 && p->w_packed && p->M > 0 && p->N > 0 && p->k0 > 0 && p->k1 >= 0, MMG_EINVAL, "gen_layer_fwd: bad arguments");
    const int K = p->k0 + p->k1;
    MMG_REQUIRE(K <= GT_MAX_K, MMG_EUNSUPPORTED, "gen_layer_fwd: at most %d input features (got %d)", GT_MAX_K, K);
    MMG_REQUIRE(p->in_mode >= 0 && p->in_mode <= 2, MMG_EINVAL, "gen_layer_fwd: bad in_mode");
    if (p->in_mode != 0) {
        MMG_REQUIRE(p->k1 == 0 && (K & 7) == 0 && ((uintptr_t)p->x0 & 15) == 0, MMG_EUNSUPPORTED, "gen_layer_fwd: BN prologue needs one 16-byte-aligned input with K % 8 == 0");
        MMG_REQUIRE(p->in_gamma && p->in_beta, MMG_EINVAL, "gen_layer_fwd: missing input BN affine");
        MMG_REQUIRE(p->in_mode == 1 ? p->in_sums != nullptr : (p->in_run_mean && p->in_run_var), MMG_EINVAL, "gen_layer_fwd: missing input BN statistics");
        MMG_REQUIRE(p->in_mode != 1 || (p->stat_count > 0 ? p->stat_count : p->M) > 1, MMG_EINVAL, "Expected more than 1 value per channel when training, got input size (%lld, %d)", (long long)p->M, K);
    } else {
        MMG_REQUIRE(p->k1 == 0 || p->x1, MMG_EINVAL, "gen_layer_fwd: missing second input");
    }
    if (p->y_out) {
        MMG_REQUIRE(p->out_gamma && p->out_beta && (p->out_mode == 1 ? p->y_sums != nullptr : (p->out_mode == 2 && p->out_run_mean && p->out_run_var)), MMG_EINVAL,
                    "gen_layer_fwd: missing output BN parameters");
        MMG_REQUIRE(p->out_mode != 1 || (p->stat_count > 0 ? p->stat_count : p->M) > 1, MMG_EINVAL, "Expected more than 1 value per channel when training, got input size (%lld, %d)", (long long)p->M, p->N);
    }
    GenDev a;
    a.x0 = p->x0; a.x1 = p->x1; a.k0 = p->k0; a.k1 = p->k1; a.in_mode = p->in_mode;
    a.in_sums = p->in_sums; a.in_gamma = p->in_gamma; a.in_beta = p->in_beta; a.in_run_mean = p->in_run_mean; a.in_run_var = p->in_run_var;
    a.bias = p->bias; a.N = p->N; a.K = K; a.Kp = round_up(K, 64);
    a.NG = p->N > 256 ? 256 : round_up(p->N, 32);
    a.n_groups = (p->N + a.NG - 1) / a.NG;
    a.z_out = p->z_out; a.out_sums = p->out_sums; a.y_out = p->y_out; a.out_mode = p->out_mode; a.y_sums = p->y_sums;
    a.out_gamma = p->out_gamma; a.out_beta = p->out_beta; a.out_run_mean = p->out_run_mean; a.out_run_var = p->out_run_var;
    a.momentum = p->momentum; a.eps = p->eps; a.update_running = p->update_running; a.M = p->M;
    a.cnt = (double)(p->stat_count > 0 ? p->stat_count : p->M);
    const long long row_tiles = (p->M + 127) / 128;
    MMG_REQUIRE(row_tiles * a.n_groups < (1LL << 30), MMG_EUNSUPPORTED, "gen_layer_fwd: batch too large");
    a.row_tiles = (int)row_tiles;
    const int Np = a.NG * a.n_groups;
    CUtensorMap map_w;
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w, p->w_packed, (uint64_t)a.Kp, (uint64_t)Np, (uint64_t)a.Kp * 2, 64, (uint32_t)a.NG, CU_TENSOR_MAP_SWIZZLE_128B) == 0,
                MMG_EINVAL, "gen_layer_fwd: cuTensorMapEncodeTiled(w) failed");
    const int kchunks = a.Kp / 64;
    const int wg = g_worker_groups;
    const long long items = row_tiles * a.n_groups;
    const int grid = (int)(items < MMG_NUM_SMS ? items : MMG_NUM_SMS);
    const bool single = items <= grid;                       // one item per CTA: one weight stage, one accumulator
    a.w_stages = single ? 1 : GT_W_STAGES;
    a.acc_stride = single ? 0 : a.NG;
    int cols = single ? a.NG : 2 * a.NG;
    a.tm_cols = 32;
    while (a.tm_cols < cols) a.tm_cols *= 2;
    size_t smem = 1024 + (size_t)kchunks * GT_A_CHUNK + (size_t)a.w_stages * kchunks * a.NG * 128 + (p->y_out ? 4 * GT_Y_STAGE : 0) +
                  (p->out_sums ? (size_t)8 * a.NG * sizeof(float) : 0);
    // a pure y pass (the wide output layer) keeps scale / shift of all its columns in shared memory when they fit
    const size_t cache = (size_t)2 * a.n_groups * a.NG * sizeof(float);
    a.col_cache = p->y_out && !p->z_out && !p->out_sums && smem + cache <= 220 * 1024;
    if (a.col_cache) smem += cache;
    CUtensorMap map_y = map_w;
    if (p->y_out) {
        MMG_REQUIRE(((uintptr_t)p->y_out & 15) == 0 && (p->N & 3) == 0, MMG_EUNSUPPORTED, "gen_layer_fwd: y_out must be 16-byte aligned with N % 4 == 0");
        MMG_REQUIRE(tc::make_map_2d(&map_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p->y_out, (uint64_t)p->N, (uint64_t)p->M, (uint64_t)p->N * 4, 32, 128,
                                    CU_TENSOR_MAP_SWIZZLE_128B) == 0, MMG_EINVAL, "gen_layer_fwd: cuTensorMapEncodeTiled(y) failed");
    }
    MMG_REQUIRE(smem <= 220 * 1024, MMG_EUNSUPPORTED, "gen_layer_fwd: tile does not fit shared memory");
    static bool attr_done = false;
    if (!attr_done) {
        MMG_CUDA(cudaFuncSetAttribute(gen_layer_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        MMG_CUDA(cudaFuncSetAttribute(gen_layer_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done = true;
    }
    // (TMEM columns and shared memory are sized to need, so single-item launches of the two generators COULD share an SM; a 2-CTAs-per-SM build
    // of the kernel (96 registers) was measured with the beat generator on its side stream: 224 us for both forwards instead of 213 -- dropped)
    if (wg == 4) gen_layer_tc_kernel<4><<<grid, 64 + 4 * 128, smem, (cudaStream_t)stream>>>(map_w, map_y, a);
    else gen_layer_tc_kernel<2><<<grid, 64 + 2 * 128, smem, (cudaStream_t)stream>>>(map_w, map_y, a);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
