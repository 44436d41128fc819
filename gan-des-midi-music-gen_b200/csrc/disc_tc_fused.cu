// Fused backward of the bf16 tensor-core discriminator (autograd of DiscriminatorCNN, network_tests.py:147-160, reached by
// disc_loss.backward() / gen_loss.backward(), network_tests.py:307,314): ONE persistent kernel per pass instead of
// fc_bwd + conv2_wgrad + conv2_dgrad + conv1_wgrad.  Layouts: see disc_tc.cu.
//
// A CTA walks over whole samples.  Per sample it streams in what the forward left in HBM -- XS (27 KB), P1 (55 KB), A2
// (27 KB), all bf16 -- and everything else lives on chip:
//   W1  workers   A2 -> DZ2 in place in shared memory (dz2 = dlogit * fc.w * lrelu'(a2)); dfc.w accumulators are
//                 read-modify-written in TMEM (tcgen05.ld / tcgen05.st), dconv2.b in registers
//   M1  tcgen05   conv2 weight gradient (P1 x DZ2, both MN-major, accumulator persistent in TMEM) and conv2 data gradient
//                 (DZ2 tap-shift x W2d -> 4 x 64 TMEM columns)
//   W3  workers   dgrad epilogue: x lrelu'(a1) (mask from P1 in shared memory) -> DZ1C rows in shared memory; dconv1.b in registers
//   M2  tcgen05   conv1 weight gradient (XS x DZ1C, accumulator persistent in TMEM)
// so DZ2 and DZ1C (82 KB per sample) never touch HBM and the weight gradients leave the SM once per launch.
// The TMA producer prefetches sample i+1's A2 / P1 / XS as soon as the MMAs / the epilogue of sample i release the buffers.
// HBM per sample and pass: 110 KB read (the unfused chain: 382 KB read + 137 KB written).
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace {

constexpr int P1_ROWS = 429, P1_W = 13;       // 33 x 13 super pixels per sample (conv1 activations / conv2 row space)
constexpr int XS_ROWS = 1690, XS_W = 26;      // 65 x 26 super pixels per sample (input / conv1 row space)

constexpr int FB_WORKERS = 256;
constexpr int FB_THREADS = 64 + FB_WORKERS;   // warp 0 TMA, warp 1 MMA, warps 2-9 workers

// shared-memory map (offsets from a 1024-byte aligned base)
constexpr int SM_W2D = 0;                     // [4 t][64 n][32 oc] bf16, SW64                16384
constexpr int SM_P1 = SM_W2D + 16384;         // 448 rows x 128 B, SW128                      57344
constexpr int SM_DZ2 = SM_P1 + 57344;         // 16 zero halo rows + 512 rows, 64 B, SW64     33792
constexpr int SM_XS = SM_DZ2 + 33792;         // 1792 rows x 16 B, no swizzle                 28672
constexpr int SM_DZ1 = SM_XS + 28672;         // 1696 rows x 32 B, SW32                       54272
constexpr int SM_TOTAL = SM_DZ1 + 54272;      // 190464
constexpr int P1_LOAD_ROWS = 448, A2_LOAD_ROWS = 432, XS_LOAD_ROWS = 1792;   // 2 x 224, 2 x 216, 7 x 256 row boxes
constexpr int K2_STEPS = 27;                  // conv2 wgrad: 27 x 16 = 432 rows
constexpr int K1_STEPS = 106;                 // conv1 wgrad: 106 x 16 = 1696 rows

// TMEM columns
constexpr uint32_t TM_W2 = 0;                 // 64:  conv2.weight gradient  [ty][oc]
constexpr uint32_t TM_W1 = 64;                // 32:  conv1.weight gradient  [ty][oc]
constexpr uint32_t TM_FC = 128;               // 128: fc.weight gradient     [tile][oc]
constexpr uint32_t TM_DG = 256;               // 256: conv2 dgrad            [tile][cell*16 + ic]

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                   "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct FusedBwdArgs {
    const float* dlogit; const float* wfcp;
    float* dw1; float* db1; float* dw2; float* db2; float* dwfc; float* dbfc;
    int B; int dbg_skip;      // dbg_skip (env MMG_DBG_SKIP, timing experiments only): bit0 conv2 wgrad, bit1 conv2 dgrad, bit2 conv1 wgrad MMAs are not issued
};

__global__ void __launch_bounds__(FB_THREADS, 1) disc_bwd_fused_kernel(const __grid_constant__ CUtensorMap map_xs, const __grid_constant__ CUtensorMap map_p1,
                                                                       const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_w,
                                                                       const FusedBwdArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_p1, full_a2, full_xs, empty_p1, dz2_ready, dz1_ready, mma1_done, mma2_done, wbar;
    __shared__ uint32_t tmem_s;
    __shared__ float red_s[48];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = a.B > (int)blockIdx.x ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;      // samples blockIdx.x, +gridDim.x, ...

    if (threadIdx.x == 0) {
        tc::mbar_init(&full_p1, 1); tc::mbar_init(&full_a2, 1); tc::mbar_init(&full_xs, 1); tc::mbar_init(&wbar, 1);
        tc::mbar_init(&empty_p1, FB_WORKERS); tc::mbar_init(&dz2_ready, FB_WORKERS); tc::mbar_init(&dz1_ready, FB_WORKERS);
        tc::mbar_init(&mma1_done, 1); tc::mbar_init(&mma2_done, 1);
        tc::fence_barrier_init();
    }
    if (threadIdx.x < 48) red_s[threadIdx.x] = 0.f;
    if (warp == 1) { tc::tmem_alloc(&tmem_s, 512); tc::tmem_relinquish(); }
    {   // DZ2 (halo + tail rows), XS and DZ1C start as zeros: rows the per-sample passes never write must read as 0
        uint4* z = reinterpret_cast<uint4*>(smem + SM_DZ2);
        for (int i = threadIdx.x; i < (SM_TOTAL - SM_DZ2) / 16; i += FB_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0 && n_my > 0) {
            tc::mbar_expect_tx(&wbar, 16384);
            tc::tma_load_2d(smem + SM_W2D, &map_w, &wbar, 0, 0);
            for (int it = 0; it < n_my; ++it) {
                const int b = blockIdx.x + it * gridDim.x;
                const uint32_t prev = (uint32_t)((it - 1) & 1);
                if (it > 0) tc::mbar_wait(&mma1_done, prev);                      // conv2 wgrad / dgrad MMAs of the previous sample have read DZ2
                tc::mbar_expect_tx(&full_a2, A2_LOAD_ROWS * 64);
                tc::tma_load_2d(smem + SM_DZ2 + 1024, &map_a2, &full_a2, 0, b * P1_ROWS);
                tc::tma_load_2d(smem + SM_DZ2 + 1024 + 216 * 64, &map_a2, &full_a2, 0, b * P1_ROWS + 216);
                if (it > 0) tc::mbar_wait(&empty_p1, prev);                       // its dgrad epilogue has read P1
                tc::mbar_expect_tx(&full_p1, P1_LOAD_ROWS * 128);
                tc::tma_load_2d(smem + SM_P1, &map_p1, &full_p1, 0, b * P1_ROWS);
                tc::tma_load_2d(smem + SM_P1 + 224 * 128, &map_p1, &full_p1, 0, b * P1_ROWS + 224);
                if (it > 0) tc::mbar_wait(&mma2_done, prev);                      // its conv1 wgrad MMAs have read XS
                tc::mbar_expect_tx(&full_xs, XS_LOAD_ROWS * 16);
#pragma unroll
                for (int j = 0; j < 7; ++j) tc::tma_load_2d(smem + SM_XS + j * 256 * 16, &map_xs, &full_xs, 0, b * XS_ROWS + j * 256);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (n_my > 0 && tc::elect_one()) {
            const uint32_t leader = 1;
            constexpr uint64_t P1_MN = tc::smem_desc_base(128, 1024, tc::SW_128B);    // conv2 wgrad A: atom 1 = one row (128 B) later
            constexpr uint64_t DZ2_MN = tc::smem_desc_base(0, 512, tc::SW_64B);       // conv2 wgrad B
            constexpr uint64_t KM64 = tc::smem_desc_base(0, 512, tc::SW_64B);         // conv2 dgrad A (DZ2 rows) and B (W2d)
            constexpr uint64_t XS_MN = tc::smem_desc_base(128, 16, tc::SW_NONE);      // conv1 wgrad A: atoms one 16-byte row apart
            constexpr uint64_t DZ1_MN = tc::smem_desc_base(0, 256, tc::SW_32B);       // conv1 wgrad B
            constexpr uint32_t ID_WG2 = tc::idesc_bf16(128, 32, 1, 1), ID_DG = tc::idesc_bf16(128, 64), ID_WG1 = tc::idesc_bf16(64, 16, 1, 1);
            const uint32_t p1 = tc::smem_u32(smem + SM_P1), dz2 = tc::smem_u32(smem + SM_DZ2) + 1024, w2d = tc::smem_u32(smem + SM_W2D);
            const uint32_t xs = tc::smem_u32(smem + SM_XS), dz1 = tc::smem_u32(smem + SM_DZ1);
            tc::mbar_wait(&wbar, 0);
            for (int it = 0; it < n_my; ++it) {
                const uint32_t ph = (uint32_t)(it & 1);
                tc::mbar_wait(&dz2_ready, ph);
                tc::mbar_wait(&full_p1, ph);
                tc::tc_fence_after();
#pragma unroll
                for (int ty = 0; ty < 2 && !(a.dbg_skip & 1); ++ty)
#pragma unroll 9
                    for (int k = 0; k < K2_STEPS; ++k)
                        tc::mma_f16_ss_pred(tmem + TM_W2 + ty * 32, tc::smem_desc(P1_MN, p1 + (ty * P1_W + k * 16) * 128), tc::smem_desc(DZ2_MN, dz2 + k * 16 * 64),
                                       ID_WG2, (it | k) != 0, leader);
#pragma unroll
                for (int tile = 0; tile < 4 && !(a.dbg_skip & 2); ++tile)
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            tc::mma_f16_ss_pred(tmem + TM_DG + tile * 64, tc::smem_desc(KM64, dz2 + (tile * 128 - ((t >> 1) * P1_W + (t & 1))) * 64 + k * 32),
                                           tc::smem_desc(KM64, w2d + t * 4096 + k * 32), ID_DG, (t | k) != 0, leader);
                tc::mma_commit_pred(&mma1_done, leader);
                tc::mbar_wait(&dz1_ready, ph);
                tc::mbar_wait(&full_xs, ph);
                tc::tc_fence_after();
#pragma unroll
                for (int ty = 0; ty < 2 && !(a.dbg_skip & 4); ++ty)
#pragma unroll 8
                    for (int k = 0; k < K1_STEPS; ++k)
                        tc::mma_f16_ss_pred(tmem + TM_W1 + ty * 16, tc::smem_desc(XS_MN, xs + (ty * XS_W + k * 16) * 16), tc::smem_desc(DZ1_MN, dz1 + k * 16 * 32),
                                       ID_WG1, (it | k) != 0, leader);
                tc::mma_commit_pred(&mma2_done, leader);
            }
        }
    } else {
        // ------------------------------------------------------------------ workers: thread (q, lane, h) owns TMEM lane q*32+lane, column half h
        const int q = warp & 3, h = (warp - 2) >> 2;
        const int tl = q * 32 + lane;                                 // row inside a 128-row tile = TMEM lane
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        float db2[16], db1[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) db2[c] = db1[c] = 0.f;
        float dbfc = 0.f;
        {   // fc.weight gradient accumulators start at zero
            uint32_t zr[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) zr[c] = 0u;
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) tmem_st_32x16(tmem + tlane + TM_FC + tile * 32 + h * 16, zr);
            tmem_st_wait();
        }
        const uint32_t dz2s = tc::smem_u32(smem + SM_DZ2) + 1024, p1s = tc::smem_u32(smem + SM_P1), dz1s = tc::smem_u32(smem + SM_DZ1);
        // fc.weight slice of this thread's rows (the same rows for every sample): loaded once, kept in registers
        float wreg[4][16];
#pragma unroll
        for (int tile = 0; tile < 4; ++tile) {
            const int R = tile * 128 + tl;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 w = R < P1_ROWS ? reinterpret_cast<const float4*>(a.wfcp + R * 32 + h * 16)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
                wreg[tile][4 * c4] = w.x; wreg[tile][4 * c4 + 1] = w.y; wreg[tile][4 * c4 + 2] = w.z; wreg[tile][4 * c4 + 3] = w.w;
            }
        }
        float dl_next = n_my > 0 ? a.dlogit[blockIdx.x] : 0.f;
        for (int it = 0; it < n_my; ++it) {
            const uint32_t ph = (uint32_t)(it & 1);
            const float dl = dl_next;
            if (it + 1 < n_my) dl_next = a.dlogit[blockIdx.x + (it + 1) * gridDim.x];      // in flight behind this sample's work
            if (threadIdx.x == 64) dbfc += dl;
            // ---- W1: A2 -> DZ2 in place, fc.weight / conv2.bias gradients
            tc::mbar_wait(&full_a2, ph);
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {
                const int R = tile * 128 + tl;
                if (tile * 128 + q * 32 >= A2_LOAD_ROWS) continue;    // warp-uniform: this warp's 32 rows are all beyond the loaded rows
                const bool loaded = R < A2_LOAD_ROWS;                 // rows 429..431 (the next sample's A2) have w = 0 -> they are zeroed
                const uint32_t rowp = dz2s + (loaded ? R : 0) * 64;
                const int sw = (R >> 1) & 3;
                const uint32_t c0p = rowp + (((2 * h) ^ sw) << 4), c1p = rowp + (((2 * h + 1) ^ sw) << 4);
                const uint4 av0 = tc::lds128(c0p), av1 = tc::lds128(c1p);
                const uint32_t au[8] = {av0.x, av0.y, av0.z, av0.w, av1.x, av1.y, av1.z, av1.w};
                uint32_t acc[16], o[8];
                tc::tmem_ld_32x16(tmem + tlane + TM_FC + tile * 32 + h * 16, acc);    // .sync.aligned: every lane of the warp takes part
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x0 = bf_lo(au[j]), x1 = bf_hi(au[j]);
                    const float g0 = dl * wreg[tile][2 * j] * (x0 > 0.f ? 1.f : 0.2f), g1 = dl * wreg[tile][2 * j + 1] * (x1 > 0.f ? 1.f : 0.2f);
                    o[j] = pack_bf16x2(g0, g1);
                    db2[2 * j] += bf_lo(o[j]); db2[2 * j + 1] += bf_hi(o[j]);          // what the MMAs will read
                    const bool real = R < P1_ROWS;
                    acc[2 * j] = __float_as_uint(fmaf(dl, real ? x0 : 0.f, __uint_as_float(acc[2 * j])));
                    acc[2 * j + 1] = __float_as_uint(fmaf(dl, real ? x1 : 0.f, __uint_as_float(acc[2 * j + 1])));
                }
                tmem_st_32x16(tmem + tlane + TM_FC + tile * 32 + h * 16, acc);
                if (loaded) {
                    tc::sts128(c0p, make_uint4(o[0], o[1], o[2], o[3]));
                    tc::sts128(c1p, make_uint4(o[4], o[5], o[6], o[7]));
                }
            }
            tmem_st_wait();
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&dz2_ready);
            // ---- W3: conv2 dgrad epilogue -> DZ1C rows, conv1.bias gradient
            tc::mbar_wait(&mma1_done, ph);
            if (it > 0) tc::mbar_wait(&mma2_done, (uint32_t)((it - 1) & 1));            // conv1 wgrad of the previous sample has read DZ1C
            tc::tc_fence_after();
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {
                const int R = tile * 128 + tl;
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem + tlane + TM_DG + tile * 64 + h * 32, r);      // cells (dy = h, dx = 0 | 1) x 16 channels
                tc::tmem_ld_wait();
                if (R >= P1_ROWS) continue;
                const int sy = R / P1_W, sx = R - sy * P1_W;
                const int oy = 2 * sy + h - 1;
                if (oy < 0 || oy >= 64) continue;                     // zero-padding cells of P1: no conv1 output behind them
                const uint32_t prow = p1s + R * 128;
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int ox = 2 * sx + dx - 1;
                    if (ox < 0 || ox >= 25) continue;
                    const int m = oy * XS_W + ox;
                    const uint32_t drow = dz1s + m * 32;
                    const int sw1 = (m >> 2) & 1;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {                  // 8 channels per 16-byte chunk
                        const int pc = (h * 2 + dx) * 2 + hh;
                        const uint4 av = tc::lds128(prow + ((pc ^ (R & 7)) << 4));
                        const uint32_t au[4] = {av.x, av.y, av.z, av.w};
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int c = dx * 16 + hh * 8 + 2 * j;
                            const float g0 = __uint_as_float(r[c]) * (bf_lo(au[j]) > 0.f ? 1.f : 0.2f);
                            const float g1 = __uint_as_float(r[c + 1]) * (bf_hi(au[j]) > 0.f ? 1.f : 0.2f);
                            o[j] = pack_bf16x2(g0, g1);
                            db1[hh * 8 + 2 * j] += bf_lo(o[j]);
                            db1[hh * 8 + 2 * j + 1] += bf_hi(o[j]);
                        }
                        tc::sts128(drow + ((hh ^ sw1) << 4), make_uint4(o[0], o[1], o[2], o[3]));
                    }
                }
            }
            tc::tc_fence_before();
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&empty_p1);
            tc::mbar_arrive(&dz1_ready);
        }
        // ------------------------------------------------------------------ flush: weight gradients leave the SM once per launch
        if (n_my > 0) {
            tc::mbar_wait(&mma2_done, (uint32_t)((n_my - 1) & 1));
            tc::tc_fence_after();
            if (h == 0) {
                // conv2: TMEM lane m = tx*64 + (dy*2+dx)*16 + ic, column = ty*32 + oc  ->  conv2.weight[oc][ic][2ty+dy][2tx+dx]
                const int tx = tl >> 6, dy = (tl >> 5) & 1, dx = (tl >> 4) & 1, ic = tl & 15;
#pragma unroll
                for (int ty = 0; ty < 2; ++ty) {
                    uint32_t r[32];
                    tc::tmem_ld_32x32(tmem + tlane + TM_W2 + ty * 32, r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int oc = 0; oc < 32; ++oc) atomicAdd(&a.dw2[((oc * 16 + ic) * 4 + 2 * ty + dy) * 4 + 2 * tx + dx], __uint_as_float(r[oc]));
                }
                if (q == 0) {
                    // conv1 (M = 64 accumulator): row i sits in TMEM lane (i % 16) + 32 * (i / 16); rows 0..15 = the two horizontal taps
#pragma unroll
                    for (int ty = 0; ty < 2; ++ty) {
                        uint32_t r[16];
                        tc::tmem_ld_32x16(tmem + TM_W1 + ty * 16, r);
                        tc::tmem_ld_wait();
                        if (lane < 16) {
                            const int tx1 = lane >> 3, e = lane & 7, dy1 = e >> 2, dx1 = (e >> 1) & 1, ch = e & 1;
#pragma unroll
                            for (int oc = 0; oc < 16; ++oc) atomicAdd(&a.dw1[((oc * 2 + ch) * 4 + 2 * ty + dy1) * 4 + 2 * tx1 + dx1], __uint_as_float(r[oc]));
                        }
                    }
                }
            }
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {                    // fc.weight[0, oc*384 + oy*12 + ox]
                const int R = tile * 128 + tl;
                uint32_t r[16];
                tc::tmem_ld_32x16(tmem + tlane + TM_FC + tile * 32 + h * 16, r);
                tc::tmem_ld_wait();
                const int oy = R / P1_W, ox = R - oy * P1_W;
                if (R < P1_ROWS && oy < 32 && ox < 12) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) atomicAdd(&a.dwfc[(h * 16 + c) * 384 + oy * 12 + ox], __uint_as_float(r[c]));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float s2 = warp_sum(db2[c]), s1 = warp_sum(db1[c]);
            if (lane == 0) { atomicAdd(&red_s[h * 16 + c], s2); atomicAdd(&red_s[32 + c], s1); }
        }
        if (threadIdx.x == 64 && dbfc != 0.f) atomicAdd(a.dbfc, dbfc);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 512);
    if (threadIdx.x < 32) atomicAdd(&a.db2[threadIdx.x], red_s[threadIdx.x]);
    else if (threadIdx.x < 48) atomicAdd(&a.db1[threadIdx.x - 32], red_s[threadIdx.x]);
}

}  // namespace

extern "C" {

// xs (B*1690,8), p1 (B*429,64), a2 (B*429,32) bf16 as left by the forward; dlogit (B,) fp32.  All six gradients are ACCUMULATED (+=)
// into fp32 tensors in the reference's parameter layouts.  a2 is read-only here (DZ2 never leaves the SM).
int mmg_disc_bwd_fused(const void* xs, const void* p1, const void* a2, const float* dlogit, const void* packed, float* dconv1_w, float* dconv1_b,
                       float* dconv2_w, float* dconv2_b, float* dfc_w, float* dfc_b, int64_t B, void* stream) {
    MMG_REQUIRE(xs && p1 && a2 && dlogit && packed && dconv1_w && dconv1_b && dconv2_w && dconv2_b && dfc_w && dfc_b && B >= 0, MMG_EINVAL,
                "disc_bwd_fused: bad arguments");
    if (B == 0) return MMG_OK;
    MMG_REQUIRE(B * XS_ROWS < (1LL << 31) - 4096, MMG_EUNSUPPORTED, "disc_bwd_fused: batch too large");
    const unsigned char* pk = (const unsigned char*)packed;
    CUtensorMap map_xs, map_p1, map_a2, map_w;
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_xs, xs, 8, (uint64_t)(B * XS_ROWS), 16, 8, 256, CU_TENSOR_MAP_SWIZZLE_NONE) == 0, MMG_EINVAL, "disc_bwd_fused: tensor map (xs)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_p1, p1, 64, (uint64_t)(B * P1_ROWS), 128, 64, 224, CU_TENSOR_MAP_SWIZZLE_128B) == 0, MMG_EINVAL, "disc_bwd_fused: tensor map (p1)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_a2, a2, 32, (uint64_t)(B * P1_ROWS), 64, 32, 216, CU_TENSOR_MAP_SWIZZLE_64B) == 0, MMG_EINVAL, "disc_bwd_fused: tensor map (a2)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w, pk + 2048 + 16384, 32, 256, 64, 32, 256, CU_TENSOR_MAP_SWIZZLE_64B) == 0, MMG_EINVAL, "disc_bwd_fused: tensor map (w2d)");
    FusedBwdArgs a;
    a.dlogit = dlogit; a.wfcp = (const float*)(pk + 2048 + 32768);
    a.dw1 = dconv1_w; a.db1 = dconv1_b; a.dw2 = dconv2_w; a.db2 = dconv2_b; a.dwfc = dfc_w; a.dbfc = dfc_b; a.B = (int)B;
    { const char* e = getenv("MMG_DBG_SKIP"); a.dbg_skip = e ? atoi(e) : 0; }
    const int grid = (int)(B < MMG_NUM_SMS ? B : MMG_NUM_SMS);
    MMG_CUDA(cudaFuncSetAttribute(disc_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL + 1024));
    disc_bwd_fused_kernel<<<grid, FB_THREADS, SM_TOTAL + 1024, (cudaStream_t)stream>>>(map_xs, map_p1, map_a2, map_w, a);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
