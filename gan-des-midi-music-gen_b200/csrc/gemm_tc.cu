// Tensor-core GEMM for the GAN-DES contractions (GAN_DES/SIMNN.py:70-84 ConvTranspose stack, :123-142 conv2 / fc1 / fc2) and the data
// movement around it.  One kernel:
//
//     C[m][n] (+)= sum_k A(m,k) * B(n,k)          fp32 accumulators in TMEM, bf16 (kind::f16) or fp32-as-tf32 (kind::tf32) operands
//
// with each operand either K-major (stored [rows = M or N][K], K contiguous) or MN-major (stored [K][M or N], bf16 only), so that every
// layer's forward / data-gradient / weight-gradient contraction reads its tensors where they lie:
//
//   fc1 forward      y^T[o][b]   = W[o][f] x[b][f]             A = W  (K-major, M = 128 outputs), B = x (K-major), split-K over f = 55 296
//   fc1 dgrad        dx^T[f][b]  = W[o][f] dz[b][o]            A = W  (MN-major, M = f), B = dz (K-major, K = o)
//   fc1 wgrad        dW[o][f]    = dz[b][o] x[b][f]            A = dz (MN-major), B = x (MN-major), K = batch
//   conv forward     y[p][oc]    = col[p][k] W[oc][k]          A = im2col(x), B = W
//   conv dgrad       dx[p][ci]   = col(dy)[p][k'] Wf[ci][k']   (a forward convolution with the flipped, transposed weights)
//   conv wgrad       dW^T[k][oc] = col[p][k] dyT[oc][p]        A = col (MN-major), B = dy as [oc][b*p] (K-major)
//   convT forward    col[p][co*taps] = x[p][ci] W[ci][co*taps] A = x (NHWC, K-major), B = W (MN-major), then col2im
//
// TMA (cp.async.bulk.tensor, 128-byte swizzle, zero fill outside the tensor = all M / N / K tails) feeds a 4-stage shared-memory ring;
// warp 0 = producer, warp 1 = tcgen05.mma issuer + TMEM owner, warps 2-5 = epilogue (bias, activation, transposed / NCHW store or fp32
// atomics for split-K).  The problems are small and HBM- or latency-bound (B = 30): the kernel is built for coverage of all layouts,
// not for the last percent of the tensor pipe.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int GM_STAGES = 4;
constexpr int GM_THREADS = 192;
constexpr int GM_A_BYTES = 128 * 128;      // 128 rows x one 128-byte K chunk (K-major) = two 64 x 64 MN-major boxes

struct GemmArgs {
    int M, N, BN;                          // BN = tile width (32 / 64 / 128 / 256)
    int a_mn, b_mn, tf32;
    int chunks, chunks_per_split;          // 128-byte K chunks: 64 bf16 or 32 tf32 elements
    int tiles_m, tiles;                    // persistent mode (split_k == 1): a CTA walks over output tiles blockIdx.x, + gridDim.x, ... (tile = tn * tiles_m + tm)
    int stages;                            // depth of the shared-memory ring: min(GM_STAGES, chunks per CTA) -- short-K problems then fit several CTAs per SM
    float* C;
    long long ldc;
    int trans_out;                         // element (m, n) -> C[(m / inner) * N * inner + n * inner + m % inner]   (NCHW: inner = pixels per image)
    long long inner;
    int atomic;                            // fp32 atomicAdd (split-K); bias / act are then left to the caller
    int out_bf16;                          // C is a bf16 row-major matrix (plain stores only)
    const float* bias;
    int bias_on_m;
    int act;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void mma_tf32_ss_pred(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate, uint32_t issue) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue) : "memory");
}

__global__ void __launch_bounds__(GM_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full[GM_STAGES], empty[GM_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_s;
    __shared__ __align__(16) unsigned char ep_stage[4][32 * 80];       // bf16 epilogue: 32 rows x 64 B per epilogue warp, 80-byte pitch (16-byte accesses without bank conflicts)
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // two launch shapes: split-K (grid = tiles_m x tiles_n x splits, one tile and one K range per CTA) and persistent (grid.x CTAs walk over
    // all tiles with two TMEM accumulators: the epilogue of tile t runs under the loads and MMAs of tile t + 1)
    const bool persistent = gridDim.z == 1 && gridDim.y == 1;
    const int c0 = persistent ? 0 : blockIdx.z * a.chunks_per_split;
    const int nck = min(a.chunks, c0 + a.chunks_per_split) - c0;
    if (nck <= 0) return;                                                     // (split-K tail with nothing to add)
    const int tile0 = persistent ? (int)blockIdx.x : (int)(blockIdx.y * a.tiles_m + blockIdx.x);
    const int tile_step = persistent ? (int)gridDim.x : a.tiles;              // split-K: exactly one tile
    const int tile_end = persistent ? a.tiles : tile0 + 1;
    const int b_bytes = a.BN * 128, stage_bytes = GM_A_BYTES + b_bytes;
    const int kelems = a.tf32 ? 32 : 64;

    if (threadIdx.x == 0) {
        for (int i = 0; i < GM_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 4); }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&map_a);
        tc::tma_prefetch_desc(&map_b);
    }
    const uint32_t acc_cols = a.BN < 32 ? 32u : (uint32_t)a.BN;
    const uint32_t tm_cols = persistent ? 2u * acc_cols : acc_cols;          // persistent: two accumulators (BN <= 256 -> at most all 512 columns)
    if (warp == 1) { tc::tmem_alloc(&tmem_s, tm_cols); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp == 0) {
        if (lane == 0) {
            int i = 0;                                                        // chunk counter across tiles: the ring keeps rolling
            for (int tile = tile0; tile < tile_end; tile += tile_step)
            for (int ic = 0; ic < nck; ++ic, ++i) {
                const int m0 = (tile % a.tiles_m) * 128, n0 = (tile / a.tiles_m) * a.BN;
                const int stage = i % a.stages;
                const uint32_t phase = (uint32_t)(i / a.stages) & 1u;
                tc::mbar_wait(&empty[stage], phase ^ 1u);
                tc::mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
                unsigned char* sa = smem + stage * stage_bytes;
                unsigned char* sb = sa + GM_A_BYTES;
                const int kc = (c0 + ic) * kelems;
                if (!a.a_mn) tc::tma_load_2d(sa, &map_a, &full[stage], kc, m0);                   // box (K chunk, 128 rows)
                else {                                                                              // two boxes (64 m, 64 k rows)
                    tc::tma_load_2d(sa, &map_a, &full[stage], m0, kc);
                    tc::tma_load_2d(sa + 8192, &map_a, &full[stage], m0 + 64, kc);
                }
                if (!a.b_mn) tc::tma_load_2d(sb, &map_b, &full[stage], kc, n0);                   // box (K chunk, BN rows)
                else
                    for (int j = 0; j < a.BN / 64; ++j) tc::tma_load_2d(sb + j * 8192, &map_b, &full[stage], n0 + 64 * j, kc);
            }
        }
    } else if (warp == 1) {
        const uint32_t leader = tc::elect_one() ? 1u : 0u;
        constexpr uint64_t KM128 = tc::smem_desc_base(0, 1024, tc::SW_128B);                       // K-major: 8-row groups 1024 B apart
        constexpr uint64_t MN128 = tc::smem_desc_base(8192, 1024, tc::SW_128B);                    // MN-major: 64-element MN atoms 8192 B apart, 8-row K groups 1024 B apart
        const uint32_t idesc = a.tf32 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.BN >> 3) << 17) | (8u << 24))
                                      : tc::idesc_bf16(128, (uint32_t)a.BN, (uint32_t)a.a_mn, (uint32_t)a.b_mn);
        const uint64_t da = a.a_mn ? MN128 : KM128, db = a.b_mn ? MN128 : KM128;
        const uint32_t sa_step = a.a_mn ? 2048u : 32u, sb_step = a.b_mn ? 2048u : 32u;            // one MMA = 16 bf16 / 8 tf32 K elements = 32 B, or 16 K rows
        int i = 0, it = 0;
        for (int tile = tile0; tile < tile_end; tile += tile_step, ++it) {
        const int buf = it & 1;
        const uint32_t tacc = tmem + (uint32_t)buf * acc_cols;
        tc::mbar_wait(&acc_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);        // the epilogue has drained this accumulator (first two uses: free)
        tc::tc_fence_after();
        for (int ic = 0; ic < nck; ++ic, ++i) {
            const int stage = i % a.stages;
            const uint32_t phase = (uint32_t)(i / a.stages) & 1u;
            tc::mbar_wait(&full[stage], phase);
            tc::tc_fence_after();
            const uint32_t sa = tc::smem_u32(smem + stage * stage_bytes), sb = sa + GM_A_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (a.tf32) mma_tf32_ss_pred(tacc, tc::smem_desc(da, sa + k * sa_step), tc::smem_desc(db, sb + k * sb_step), idesc, (ic | k) != 0, leader);
                else tc::mma_f16_ss_pred(tacc, tc::smem_desc(da, sa + k * sa_step), tc::smem_desc(db, sb + k * sb_step), idesc, (ic | k) != 0, leader);
            }
            tc::mma_commit_pred(&empty[stage], leader);
        }
        tc::mma_commit_pred(&acc_full[buf], leader);
        }
    } else {
        const int q = warp & 3;                                                                     // TMEM lane quadrant of this warp
        int it = 0;
        for (int tile = tile0; tile < tile_end; tile += tile_step, ++it) {
        const int buf = it & 1;
        const uint32_t tacc = tmem + (uint32_t)buf * acc_cols;
        const int m0 = (tile % a.tiles_m) * 128, n0 = (tile / a.tiles_m) * a.BN;
        const int m = m0 + q * 32 + lane;
        tc::mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
        tc::tc_fence_after();
        const long long mo = a.trans_out ? (long long)(m / a.inner) * (long long)a.N * a.inner + (m % a.inner) : (long long)m * a.ldc;
        const float bm = (a.bias && a.bias_on_m && m < a.M) ? a.bias[m] : 0.f;
        for (int c = 0; c < a.BN; c += 32) {
            uint32_t r[32];
            tc::tmem_ld_32x32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
            tc::tmem_ld_wait();
            if (a.out_bf16) {
                // bf16 row-major output (the tap columns of a convolution data gradient).  A lane holds 32 columns of ITS row (64 B); stored
                // directly, every instruction would scatter 32 half-sector pieces over 32 rows (110 MB went at 0.9 TB/s).  The warp's 32 x 64 B tile
                // is transposed through shared memory instead: four lanes then write one row's 64 B, eight rows per instruction, whole sectors.
                __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(a.C);
                if ((n0 + c + 32 <= a.N || ((a.N & 7) == 0 && n0 + c < a.N)) && (a.ldc & 7) == 0 && ((uintptr_t)a.C & 15) == 0) {      // (a tail of whole 8-column pieces too)
                    unsigned char* st = ep_stage[warp - 2];
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])); o.y = pack_bf16x2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                        o.z = pack_bf16x2(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5])); o.w = pack_bf16x2(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
                        *reinterpret_cast<uint4*>(st + lane * 80 + (j >> 3) * 16) = o;
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = 8 * i + (lane >> 2), piece = lane & 3;
                        const uint4 o = *reinterpret_cast<const uint4*>(st + row * 80 + piece * 16);
                        const int mr = m0 + q * 32 + row;
                        if (mr < a.M && n0 + c + piece * 8 + 8 <= a.N) *reinterpret_cast<uint4*>(cb + (long long)mr * a.ldc + n0 + c + piece * 8) = o;
                    }
                    __syncwarp();
                } else if (m < a.M && n0 + c < a.N) {
                    cb += (long long)m * a.ldc + n0 + c;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + c + j < a.N) cb[j] = __float2bfloat16(__uint_as_float(r[j]));
                }
            } else if (m < a.M && !a.trans_out && !a.atomic && n0 + c + 32 <= a.N && (a.ldc & 3) == 0 && ((uintptr_t)a.C & 15) == 0 && ((n0 + c) & 3) == 0) {
                // plain row-major store of 32 consecutive columns: eight 16-byte stores
                float4* dst = reinterpret_cast<float4*>(a.C + mo + n0 + c);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[e] = __uint_as_float(r[j + e]) + (a.bias ? (a.bias_on_m ? bm : a.bias[n0 + c + j + e]) : 0.f);
                        v[e] = mmg_act(v[e], a.act);
                    }
                    dst[j >> 2] = make_float4(v[0], v[1], v[2], v[3]);
                }
            } else if (m < a.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + c + j;
                    if (n < a.N) {
                        float v = __uint_as_float(r[j]);
                        float* dst = a.C + (a.trans_out ? mo + (long long)n * a.inner : mo + n);
                        if (a.atomic) atomicAdd(dst, v);
                        else {
                            v += a.bias ? (a.bias_on_m ? bm : a.bias[n]) : 0.f;
                            *dst = mmg_act(v, a.act);
                        }
                    }
                }
            }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);                   // this warp has read its lanes of the accumulator
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, tm_cols);
}

// ---------------------------------------------------------------------------------------------------------------- data movement
// fp32 -> bf16 with a permutation of a (d0, d1, d2) tensor: dst[i_a * a_stride + i_b * pitch + i_c] where (a, b, c) = perm of (0, 1, 2);
// columns [n_c, pitch) are written as zeros (TMA needs 16-byte row pitches; K tails must read as zeros).  One thread = 8 consecutive
// columns of one row (one 16-byte store when the row is aligned), 32-bit index arithmetic.
__global__ void pack_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int d0, int d1, int d2, int pa, int pb, int pc, long long pitch,
                                 long long a_stride) {
    auto dim = [&](int i) { return i == 0 ? d0 : i == 1 ? d1 : d2; };
    auto str = [&](int i) { return i == 0 ? (long long)d1 * d2 : i == 1 ? (long long)d2 : 1LL; };
    const unsigned nb = (unsigned)dim(pb), nc = (unsigned)dim(pc);
    const unsigned groups = (unsigned)((pitch + 7) >> 3);
    const unsigned total = (unsigned)dim(pa) * nb * groups;
    const long long sa = str(pa), sb = str(pb), sc = str(pc);
    const bool vec = (pitch & 7) == 0 && (a_stride & 7) == 0 && ((uintptr_t)dst & 15) == 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned gq = i % groups, row = i / groups, bb = row % nb, aa = row / nb;
        const unsigned c0 = gq * 8;
        const float* sp = src + aa * sa + bb * sb + (long long)c0 * sc;
        __nv_bfloat16* dp = dst + aa * a_stride + bb * pitch + c0;
        float v[8];
        if (sc == 1 && c0 + 8 <= nc && ((uintptr_t)sp & 15) == 0) {
            const float4 q0 = reinterpret_cast<const float4*>(sp)[0], q1 = reinterpret_cast<const float4*>(sp)[1];
            v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = c0 + e < nc ? sp[e * sc] : 0.f;
        }
        if (vec) {
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(dp) = o;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (c0 + e < (unsigned)pitch) dp[e] = __float2bfloat16(v[e]);
        }
    }
}

// im2col of an NCHW fp32 tensor into bf16 rows [b * OH * OW][Ci * kh * kw (+ optional ones column, + zero padding up to pitch)].
// One thread = 8 consecutive columns of one row (one 16-byte store); the (c, ky, kx) decode advances incrementally.
__global__ void im2col_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int B, int Ci, int H, int W, int kh, int kw, int stride, int pad,
                                   int OH, int OW, int pitch, int ones_col) {
    const int K = Ci * kh * kw;
    const unsigned groups = (unsigned)pitch >> 3;
    const unsigned total = (unsigned)B * OH * OW * groups;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned gq = i % groups, p = i / groups;
        const int ox = (int)(p % (unsigned)OW), t = (int)(p / (unsigned)OW), oy = t % OH, b = t / OH;
        int k = (int)gq * 8;
        int kx = k % kw, t2 = k / kw, ky = t2 % kh, c = t2 / kh;
        const float* xb = x + (long long)b * Ci * H * W;
        const int iy0 = oy * stride - pad, ix0 = ox * stride - pad;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e, ++k) {
            float val = (ones_col && k == K) ? 1.f : 0.f;                  // column K = 1: the bias rides the forward GEMM, its gradient the weight-gradient GEMM
            if (k < K) {
                const int iy = iy0 + ky, ix = ix0 + kx;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = xb[((long long)c * H + iy) * W + ix];
            }
            v[e] = val;
            if (++kx == kw) { kx = 0; if (++ky == kh) { ky = 0; ++c; } }
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(col + (long long)p * pitch + gq * 8) = o;
    }
}

// The same rows with the lanes of a warp on CONSECUTIVE PIXELS of one column group: every load instruction reads 32 consecutive floats of an
// image row (the version above gathers 4-byte elements from Ci * kh different rows per instruction), and a thread stores G = 16 columns =
// one full 32-byte sector (G = 8 when the pitch is not a multiple of 16).
template <int G>
__global__ void im2col_bf16_px_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int B, int Ci, int H, int W, int kh, int kw, int stride, int pad,
                                      int OH, int OW, int pitch, int ones_col) {
    const int K = Ci * kh * kw;
    const unsigned P = (unsigned)B * OH * OW, groups = (unsigned)pitch / G;
    const unsigned total = P * groups;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned p = i % P, gq = i / P;
        const int ox = (int)(p % (unsigned)OW), t = (int)(p / (unsigned)OW), oy = t % OH, b = t / OH;
        int k = (int)gq * G;
        int kx = k % kw, t2 = k / kw, ky = t2 % kh, c = t2 / kh;
        const float* xb = x + (long long)b * Ci * H * W;
        const int iy0 = oy * stride - pad, ix0 = ox * stride - pad;
        float v[G];
#pragma unroll
        for (int e = 0; e < G; ++e, ++k) {
            float val = (ones_col && k == K) ? 1.f : 0.f;
            if (k < K) {
                const int iy = iy0 + ky, ix = ix0 + kx;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = xb[((long long)c * H + iy) * W + ix];
            }
            v[e] = val;
            if (++kx == kw) { kx = 0; if (++ky == kh) { ky = 0; ++c; } }
        }
        uint4* dst = reinterpret_cast<uint4*>(col + (long long)p * pitch + gq * G);
#pragma unroll
        for (int q = 0; q < G / 8; ++q) {
            uint4 o;
            o.x = pack_bf16x2(v[8 * q], v[8 * q + 1]); o.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
            o.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]); o.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
            dst[q] = o;
        }
    }
}

// col2im (gather form, no atomics) of the transposed convolution: col fp32 [b * Hin * Win][Co * kh * kw (pitch ldc)] -> y NCHW [b][Co][Hout][Wout],
// y[b][co][oy][ox] = act( sum over (ky, kx) with (oy + pad - ky) % stride == 0 ... of col[b, iy, ix][co, ky, kx] )
__global__ void col2im_f32_kernel(const float* __restrict__ col, float* __restrict__ y, int B, int Co, int Hin, int Win, int kh, int kw, int stride, int pad,
                                  int Hout, int Wout, long long ldc, int act) {
    const long long total = (long long)B * Co * Hout * Wout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wout), oy = (int)((i / Wout) % Hout), co = (int)((i / ((long long)Wout * Hout)) % Co), b = (int)(i / ((long long)Wout * Hout * Co));
        float s = 0.f;
        for (int ky = 0; ky < kh; ++ky) {
            const int ty = oy + pad - ky;
            if (ty < 0 || ty % stride) continue;
            const int iy = ty / stride;
            if (iy >= Hin) continue;
            for (int kx = 0; kx < kw; ++kx) {
                const int tx = ox + pad - kx;
                if (tx < 0 || tx % stride) continue;
                const int ix = tx / stride;
                if (ix >= Win) continue;
                s += col[(((long long)b * Hin + iy) * Win + ix) * ldc + (co * kh + ky) * kw + kx];
            }
        }
        y[i] = mmg_act(s, act);
    }
}

// fp32 [rows][cols] -> fp32 transposed copy [cols][rows] of small matrices (weight-gradient layouts)
__global__ void transpose_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, long long lds) {
    const long long total = (long long)rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / rows), r = (int)(i % rows);
        dst[i] = src[(long long)r * lds + c];
    }
}

// Backward of maxpool2(relu(z)) in one pass (SIMNN.py:138-139): dz[b][c][iy][ix] = dyp[b][c][iy/2][ix/2] where (iy, ix) is the window's argmax and
// the pooled value is > 0 (the ReLU mask of the element that won), else 0; rows / columns beyond the pooled area (odd H / W) are zero.
// Outputs (either may be null): dz fp32 NCHW, and dzt bf16 [C][B*H*W] with row pitch Pp = the K-major B operand of the weight-gradient GEMM.
// One thread = 4 consecutive ix of one image row.
__global__ void pool_relu_bwd_kernel(const float* __restrict__ dyp, const uint8_t* __restrict__ idx, const float* __restrict__ yp, float* __restrict__ dz,
                                     __nv_bfloat16* __restrict__ dzt, int B, int C, int H, int W, int OH, int OW, long long Pp) {
    const unsigned WQ = (unsigned)(W + 3) >> 2;
    const unsigned total = (unsigned)B * C * H * WQ;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned wq = i % WQ, t = i / WQ, iy = t % (unsigned)H, bc = t / (unsigned)H, c = bc % (unsigned)C, b = bc / (unsigned)C;
        const unsigned oy = iy >> 1;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned ix = wq * 4 + e, ox = ix >> 1;
            v[e] = 0.f;
            if (ix < (unsigned)W && oy < (unsigned)OH && ox < (unsigned)OW) {
                const size_t o = ((size_t)bc * OH + oy) * OW + ox;
                if (idx[o] == ((iy & 1) * 2 + (ix & 1)) && yp[o] > 0.f) v[e] = dyp[o];
            }
        }
        const size_t row = ((size_t)bc * H + iy) * W + wq * 4;
        const size_t trow = (size_t)c * Pp + ((size_t)b * H + iy) * W + wq * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (wq * 4 + e < (unsigned)W) {
                if (dz) dz[row + e] = v[e];
                if (dzt) dzt[trow + e] = __float2bfloat16(v[e]);
            }
        }
    }
}

// maxpool2(relu(conv2d(x, w, b, stride 1, pad))) for a tiny stencil (SIMNN.py:123,138: Conv2d(1, 16, kernel 2) -- K = 4 is a stencil, not a
// GEMM): one thread = one pooled pixel, all output channels; the pre-pool activation never exists in memory.  Same tie-breaking and NaN
// rule as maxpool2_fwd_kernel (first maximum in (0,0), (0,1), (1,0), (1,1) order; NaN wins).
template <int CI, int KH, int KW>
__global__ void conv_small_relu_pool_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ yp,
                                            uint8_t* __restrict__ idx, int B, int H, int W, int Co, int pad, int OH2, int OW2) {
    constexpr int K = CI * KH * KW;
    __shared__ float ws[32 * K + 32];
    for (int i = threadIdx.x; i < Co * K; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < Co; i += blockDim.x) ws[32 * K + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const unsigned total = (unsigned)B * OH2 * OW2;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ox2 = (int)(i % (unsigned)OW2), t = (int)(i / (unsigned)OW2), oy2 = t % OH2, b = t / OH2;
        float patch[CI][KH + 1][KW + 1];
#pragma unroll
        for (int c = 0; c < CI; ++c)
#pragma unroll
            for (int r = 0; r <= KH; ++r)
#pragma unroll
                for (int q = 0; q <= KW; ++q) {
                    const int iy = 2 * oy2 - pad + r, ix = 2 * ox2 - pad + q;
                    patch[c][r][q] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[(((size_t)b * CI + c) * H + iy) * W + ix] : 0.f;
                }
        for (int co = 0; co < Co; ++co) {
            float v[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                float acc = ws[32 * K + co];
#pragma unroll
                for (int c = 0; c < CI; ++c)
#pragma unroll
                    for (int ky = 0; ky < KH; ++ky)
#pragma unroll
                        for (int kx = 0; kx < KW; ++kx) acc = fmaf(ws[co * K + (c * KH + ky) * KW + kx], patch[c][(d >> 1) + ky][(d & 1) + kx], acc);
                v[d] = acc > 0.f ? acc : (acc != acc ? acc : 0.f);         // ReLU (NaN propagates, as torch.relu does)
            }
            float best = v[0];
            int bi = 0;
            if (v[1] > best || v[1] != v[1]) { best = v[1]; bi = 1; }
            if (v[2] > best || v[2] != v[2]) { best = v[2]; bi = 2; }
            if (v[3] > best || v[3] != v[3]) { best = v[3]; bi = 3; }
            const size_t o = (((size_t)b * Co + co) * OH2 + oy2) * OW2 + ox2;
            yp[o] = best;
            idx[o] = (uint8_t)bi;
        }
    }
}

// Data gradient of a stride-1 convolution, second half: dcol[b * OH * OW + p][(ky * kw + kx) * Ci + ci] (bf16, the tap columns the GEMM
// dz[p][oc] x W[oc][(ky,kx,ci)] produced) -> dx[b][ci][y][x] = sum over taps of dcol[(b, y + pad - ky, x + pad - kx)][tap][ci]   (gather, no atomics).
// One thread = one input pixel: per tap one contiguous run of Ci bf16; the stores of a warp are contiguous per channel plane.
template <int CI>
__global__ void conv_dgrad_gather_kernel(const __nv_bfloat16* __restrict__ dcol, float* __restrict__ dx, int B, int H, int W, int OH, int OW, int kh, int kw, int pad,
                                         long long ldc) {
    const unsigned total = (unsigned)B * H * W;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int x = (int)(i % (unsigned)W), t = (int)(i / (unsigned)W), y = t % H, b = t / H;
        float acc[CI];
#pragma unroll
        for (int c = 0; c < CI; ++c) acc[c] = 0.f;
        for (int ky = 0; ky < kh; ++ky) {
            const int oy = y + pad - ky;
            if (oy < 0 || oy >= OH) continue;
            for (int kx = 0; kx < kw; ++kx) {
                const int ox = x + pad - kx;
                if (ox < 0 || ox >= OW) continue;
                const uint4* src = reinterpret_cast<const uint4*>(dcol + (((long long)b * OH + oy) * OW + ox) * ldc + (ky * kw + kx) * CI);
#pragma unroll
                for (int q = 0; q < CI / 8; ++q) {
                    const uint4 v = src[q];
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        acc[8 * q + 2 * e] += __uint_as_float(w4[e] << 16);
                        acc[8 * q + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CI; ++c) dx[(((size_t)b * CI + c) * H + y) * W + x] = acc[c];
    }
}

// the same gradient as bf16 NHWC rows [b * H * W][C] (the K-major A operand of the data-gradient GEMM): one thread = one pixel and 8 channels =
// one 16-byte store; the C / 8 threads of a pixel are neighbours, so a warp writes whole contiguous rows
__global__ void pool_relu_bwd_nhwc_kernel(const float* __restrict__ dyp, const uint8_t* __restrict__ idx, const float* __restrict__ yp, __nv_bfloat16* __restrict__ dzn,
                                          int B, int C, int H, int W, int OH, int OW) {
    const unsigned CG = (unsigned)C >> 3;
    const unsigned total = (unsigned)B * H * W * CG;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned cg = i % CG, p = i / CG, ix = p % (unsigned)W, t = p / (unsigned)W, iy = t % (unsigned)H, b = t / (unsigned)H;
        const unsigned oy = iy >> 1, ox = ix >> 1, code = (iy & 1) * 2 + (ix & 1);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            v[e] = 0.f;
            if (oy < (unsigned)OH && ox < (unsigned)OW) {
                const size_t o = (((size_t)b * C + cg * 8 + e) * OH + oy) * OW + ox;
                if (idx[o] == code && yp[o] > 0.f) v[e] = dyp[o];
            }
        }
        uint4 q;
        q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]); q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(dzn + (size_t)p * C + cg * 8) = q;
    }
}

// column sums of an fp32 [rows][cols] matrix (bias gradients): one warp per column group, fixed order
__global__ void colsum_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += src[(long long)r * cols + c];
    dst[c] = s;
}

// y[i][j] = act(y[i][j] + bias[j]) after a split-K accumulation
__global__ void gemm_bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, long long total, int N, int act) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        y[i] = mmg_act(y[i] + (bias ? bias[i % N] : 0.f), act);
}

}  // namespace

// A: K-major [M][K] (lda = row pitch in elements) or MN-major [K][M]; B likewise with N.  dtype 0 = bf16, 1 = fp32 read as tf32 (K-major only).
extern "C" int mmg_gemm_tc(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, float* C, long long ldc, int M, int N, int K, int dtype,
                           int split_k, int trans_out, long long inner, int atomic, const float* bias, int bias_on_m, int act, void* stream) {
    MMG_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, MMG_EINVAL, "gemm_tc: null pointer or empty problem");
    MMG_REQUIRE(dtype == 0 || (dtype == 1 && !a_mn && !b_mn), MMG_EUNSUPPORTED, "gemm_tc: tf32 operands must be K-major");
    const int esz = dtype ? 4 : 2, kelems = 128 / esz;
    MMG_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && (lda * esz) % 16 == 0 && (ldb * esz) % 16 == 0, MMG_EINVAL,
                "gemm_tc: operands need 16-byte aligned bases and row pitches");
    MMG_REQUIRE(trans_out != 1 || inner > 0, MMG_EINVAL, "gemm_tc: trans_out needs inner > 0");
    MMG_REQUIRE(!(atomic && (bias || act)), MMG_EINVAL, "gemm_tc: bias / activation cannot be fused into a split-K accumulation");
    GemmArgs a{};
    a.M = M; a.N = N;
    a.BN = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    if (b_mn && a.BN < 64) a.BN = 64;
    a.a_mn = a_mn; a.b_mn = b_mn; a.tf32 = dtype;
    a.chunks = (K + kelems - 1) / kelems;
    if (split_k < 1) split_k = 1;
    if (split_k > a.chunks) split_k = a.chunks;
    MMG_REQUIRE(split_k == 1 || atomic, MMG_EINVAL, "gemm_tc: split_k > 1 needs atomic accumulation into a zeroed C");
    a.chunks_per_split = (a.chunks + split_k - 1) / split_k;
    a.out_bf16 = trans_out == 2 ? 1 : 0;                       // trans_out: 0 = fp32 row-major, 1 = fp32 transposed / NCHW, 2 = bf16 row-major
    MMG_REQUIRE(!(a.out_bf16 && (atomic || bias || act)), MMG_EINVAL, "gemm_tc: the bf16 output takes plain stores only");
    if (a.out_bf16) trans_out = 0;
    a.C = C; a.ldc = ldc; a.trans_out = trans_out; a.inner = inner; a.atomic = atomic; a.bias = bias; a.bias_on_m = bias_on_m; a.act = act;
    CUtensorMap map_a, map_b;
    const CUtensorMapDataType dt = dtype ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    int r;
    if (!a_mn) r = tc::make_map_2d(&map_a, dt, esz, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * esz, (uint32_t)kelems, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    else r = tc::make_map_2d(&map_a, dt, esz, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * esz, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    MMG_REQUIRE(r == 0, MMG_EINVAL, "gemm_tc: cuTensorMapEncodeTiled(A) failed (%d)", r);
    if (!b_mn) r = tc::make_map_2d(&map_b, dt, esz, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * esz, (uint32_t)kelems, (uint32_t)a.BN, CU_TENSOR_MAP_SWIZZLE_128B);
    else r = tc::make_map_2d(&map_b, dt, esz, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * esz, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    MMG_REQUIRE(r == 0, MMG_EINVAL, "gemm_tc: cuTensorMapEncodeTiled(B) failed (%d)", r);
    a.tiles_m = (M + 127) / 128;
    const int tiles_n = (N + a.BN - 1) / a.BN;
    a.tiles = a.tiles_m * tiles_n;
    a.stages = split_k == 1 ? GM_STAGES : (a.chunks_per_split < GM_STAGES ? a.chunks_per_split : GM_STAGES);
    if (split_k == 1 && a.chunks < GM_STAGES && a.tiles <= 2 * MMG_NUM_SMS) a.stages = a.chunks;      // few tiles: no ring to keep rolling
    const int smem = 1024 + a.stages * (GM_A_BYTES + a.BN * 128);
    static bool attr_done = false;
    if (!attr_done) {
        MMG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + GM_STAGES * (GM_A_BYTES + 256 * 128)));
        attr_done = true;
    }
    dim3 grid((unsigned)a.tiles_m, (unsigned)tiles_n, (unsigned)split_k);
    if (split_k == 1) {                                    // persistent: as many CTAs as fit (shared memory, 512 TMEM columns), each walks over tiles
        int per_sm = (227 * 1024) / smem;
        const int by_tmem = 512 / (2 * (a.BN < 32 ? 32 : a.BN));
        if (per_sm > by_tmem) per_sm = by_tmem;
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        const int ctas = a.tiles < per_sm * MMG_NUM_SMS ? a.tiles : per_sm * MMG_NUM_SMS;
        grid = dim3((unsigned)ctas, 1, 1);
    }
    gemm_tc_kernel<<<grid, GM_THREADS, smem, (cudaStream_t)stream>>>(map_a, map_b, a);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// perm = (pa, pb, pc): destination index order over the source dimensions (d0, d1, d2); dst[i_a * a_stride + i_b * pitch + i_c] bf16, columns
// [n_pc, pitch) zero; a_stride = 0 means n_pb * pitch (dense).
extern "C" int mmg_pack_bf16(const float* src, void* dst, int d0, int d1, int d2, int pa, int pb, int pc, long long pitch, long long a_stride, void* stream) {
    MMG_REQUIRE(src && dst && d0 > 0 && d1 > 0 && d2 > 0, MMG_EINVAL, "pack_bf16: null pointer or empty tensor");
    MMG_REQUIRE(pa >= 0 && pa < 3 && pb >= 0 && pb < 3 && pc >= 0 && pc < 3 && pa != pb && pa != pc && pb != pc, MMG_EINVAL, "pack_bf16: perm must be a permutation of (0,1,2)");
    const int dims[3] = {d0, d1, d2};
    MMG_REQUIRE(pitch >= dims[pc], MMG_EINVAL, "pack_bf16: pitch smaller than the row");
    if (a_stride <= 0) a_stride = (long long)dims[pb] * pitch;
    MMG_REQUIRE(a_stride >= (long long)dims[pb] * pitch, MMG_EINVAL, "pack_bf16: a_stride smaller than one block");
    const long long total = (long long)dims[pa] * dims[pb] * ((pitch + 7) / 8);
    MMG_REQUIRE(total < (1LL << 31), MMG_EUNSUPPORTED, "pack_bf16: tensor too large for 32-bit indexing");
    pack_bf16_kernel<<<mmg_grid(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, d0, d1, d2, pa, pb, pc, pitch, a_stride);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

extern "C" int mmg_im2col_bf16(const float* x, void* col, int B, int Ci, int H, int W, int kh, int kw, int stride, int pad, int pitch, int ones_col, void* stream) {
    MMG_REQUIRE(x && col && B > 0 && Ci > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, MMG_EINVAL, "im2col_bf16: bad argument");
    const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
    MMG_REQUIRE(OH > 0 && OW > 0 && pitch >= Ci * kh * kw + (ones_col ? 1 : 0) && pitch % 8 == 0, MMG_EINVAL,
                "im2col_bf16: pitch must cover Ci*kh*kw (+1 with a ones column) and be a multiple of 8");
    MMG_REQUIRE((long long)B * OH * OW * (pitch / 8) < (1LL << 31), MMG_EUNSUPPORTED, "im2col_bf16: tensor too large for 32-bit indexing");
    if (OW >= 32) {                                         // wide images: lanes on consecutive pixels (coalesced row reads, full-sector stores)
        if (pitch % 16 == 0)
            im2col_bf16_px_kernel<16><<<mmg_grid((long long)B * OH * OW * (pitch / 16), 256, 16), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)col, B, Ci, H, W, kh, kw,
                                                                                                                                    stride, pad, OH, OW, pitch, ones_col);
        else
            im2col_bf16_px_kernel<8><<<mmg_grid((long long)B * OH * OW * (pitch / 8), 256, 16), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)col, B, Ci, H, W, kh, kw,
                                                                                                                                  stride, pad, OH, OW, pitch, ones_col);
        MMG_LAUNCH_CHECK();
        return MMG_OK;
    }
    im2col_bf16_kernel<<<mmg_grid((long long)B * OH * OW * (pitch / 8), 256, 16), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)col, B, Ci, H, W, kh, kw, stride, pad, OH, OW, pitch,
                                                                                                         ones_col);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

extern "C" int mmg_col2im_f32(const float* col, float* y, int B, int Co, int Hin, int Win, int kh, int kw, int stride, int pad, long long ldc, int act, void* stream) {
    MMG_REQUIRE(col && y && B > 0 && Co > 0 && Hin > 0 && Win > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, MMG_EINVAL, "col2im_f32: bad argument");
    const int Hout = (Hin - 1) * stride - 2 * pad + kh, Wout = (Win - 1) * stride - 2 * pad + kw;
    MMG_REQUIRE(Hout > 0 && Wout > 0 && ldc >= (long long)Co * kh * kw, MMG_EINVAL, "col2im_f32: bad geometry");
    col2im_f32_kernel<<<mmg_grid((long long)B * Co * Hout * Wout, 256), 256, 0, (cudaStream_t)stream>>>(col, y, B, Co, Hin, Win, kh, kw, stride, pad, Hout, Wout, ldc, act);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

extern "C" int mmg_transpose_f32(const float* src, float* dst, int rows, int cols, long long lds, void* stream) {
    MMG_REQUIRE(src && dst && rows > 0 && cols > 0 && lds >= cols, MMG_EINVAL, "transpose_f32: bad argument");
    transpose_f32_kernel<<<mmg_grid((long long)rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, cols, lds);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

extern "C" int mmg_colsum_f32(const float* src, float* dst, int rows, int cols, void* stream) {
    MMG_REQUIRE(src && dst && rows > 0 && cols > 0, MMG_EINVAL, "colsum_f32: bad argument");
    colsum_f32_kernel<<<(cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(src, dst, rows, cols);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

extern "C" int mmg_bias_act_inplace_f32(float* y, const float* bias, long long rows, int cols, int act, void* stream) {
    MMG_REQUIRE(y && rows > 0 && cols > 0, MMG_EINVAL, "bias_act_inplace_f32: bad argument");
    gemm_bias_act_kernel<<<mmg_grid(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(y, bias, rows * cols, cols, act);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

extern "C" int mmg_pool_relu_bwd(const float* dyp, const uint8_t* idx, const float* yp, float* dz, void* dzt, void* dzn, int B, int C, int H, int W, long long Pp,
                                 void* stream) {
    const int OH = H / 2, OW = W / 2;
    MMG_REQUIRE(dyp && idx && yp && (dz || dzt || dzn) && B > 0 && C > 0 && OH > 0 && OW > 0, MMG_EINVAL, "pool_relu_bwd: bad argument");
    MMG_REQUIRE(!dzt || Pp >= (long long)B * H * W, MMG_EINVAL, "pool_relu_bwd: pitch of the transposed output smaller than B*H*W");
    const long long total = (long long)B * C * H * ((W + 3) / 4);
    MMG_REQUIRE(total < (1LL << 31), MMG_EUNSUPPORTED, "pool_relu_bwd: tensor too large for 32-bit indexing");
    if (dz || dzt) {
        pool_relu_bwd_kernel<<<mmg_grid(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(dyp, idx, yp, dz, (__nv_bfloat16*)dzt, B, C, H, W, OH, OW, Pp);
        MMG_LAUNCH_CHECK();
    }
    if (dzn) {
        MMG_REQUIRE(C % 8 == 0 && ((uintptr_t)dzn & 15) == 0 && (long long)B * H * W * (C / 8) < (1LL << 31), MMG_EINVAL, "pool_relu_bwd: the NHWC output needs C % 8 == 0");
        pool_relu_bwd_nhwc_kernel<<<mmg_grid((long long)B * H * W * (C / 8), 256, 16), 256, 0, (cudaStream_t)stream>>>(dyp, idx, yp, (__nv_bfloat16*)dzn, B, C, H, W, OH, OW);
        MMG_LAUNCH_CHECK();
    }
    return MMG_OK;
}

/* stride-1 convolution with a tiny stencil + bias + ReLU + MaxPool2d(2,2) in one kernel; supported stencil: Ci = 1, 2 x 2 (the GAN-DES
 * discriminator's first block), Co <= 32.  yp / idx: (B, Co, OH/2, OW/2) with OH = H + 2 pad - 1. */
extern "C" int mmg_conv_small_relu_pool_f32(const float* x, const float* w, const float* bias, float* yp, uint8_t* idx, int B, int Ci, int H, int W, int Co, int kh, int kw,
                                            int pad, void* stream) {
    MMG_REQUIRE(x && w && yp && idx && B > 0 && H > 0 && W > 0 && Co > 0 && pad >= 0, MMG_EINVAL, "conv_small_relu_pool: bad argument");
    MMG_REQUIRE(Ci == 1 && kh == 2 && kw == 2 && Co <= 32, MMG_EUNSUPPORTED, "conv_small_relu_pool: only the 1 -> Co (<= 32), 2 x 2 stencil is built");
    const int OH2 = (H + 2 * pad - kh + 1) / 2, OW2 = (W + 2 * pad - kw + 1) / 2;
    MMG_REQUIRE(OH2 > 0 && OW2 > 0 && (long long)B * OH2 * OW2 < (1LL << 31), MMG_EINVAL, "conv_small_relu_pool: bad geometry");
    conv_small_relu_pool_kernel<1, 2, 2><<<mmg_grid((long long)B * OH2 * OW2, 128, 16), 128, 0, (cudaStream_t)stream>>>(x, w, bias, yp, idx, B, H, W, Co, pad, OH2, OW2);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

/* second half of the stride-1 convolution data gradient: tap columns (bf16, column (ky*kw + kx)*Ci + ci, row pitch ldc) -> dx fp32 NCHW */
extern "C" int mmg_conv_dgrad_gather(const void* dcol, float* dx, int B, int Ci, int H, int W, int kh, int kw, int pad, long long ldc, void* stream) {
    MMG_REQUIRE(dcol && dx && B > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && pad >= 0, MMG_EINVAL, "conv_dgrad_gather: bad argument");
    MMG_REQUIRE(Ci == 16, MMG_EUNSUPPORTED, "conv_dgrad_gather: built for 16 input channels (GAN-DES conv2)");
    const int OH = H + 2 * pad - kh + 1, OW = W + 2 * pad - kw + 1;
    MMG_REQUIRE(OH > 0 && OW > 0 && ldc >= (long long)kh * kw * Ci && ldc % 8 == 0 && ((uintptr_t)dcol & 15) == 0 && (long long)B * H * W < (1LL << 31), MMG_EINVAL,
                "conv_dgrad_gather: bad geometry");
    conv_dgrad_gather_kernel<16><<<mmg_grid((long long)B * H * W, 128, 16), 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dcol, dx, B, H, W, OH, OW, kh, kw, pad, ldc);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}
