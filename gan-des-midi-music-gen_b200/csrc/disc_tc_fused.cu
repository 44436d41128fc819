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
//   W3  workers   dgrad epilogue: x lrelu'(a1) -> DZ1 written IN PLACE over P1 (same super-pixel rows, same swizzle); dconv1.b in registers
//   M2  tcgen05   conv1 weight gradient in the super-pixel row space: dW[72 patch values][4 cells x 16 oc] = sum_R XS3[R][patch] * DZ1[R]
//                 (27 MMAs of M128 x N64 x K16 per sample instead of 212 of M64 x N16 x K16 in conv1's own row space).  XS3 = the 3x3
//                 neighbourhood of input super pixels of every P1 super pixel, as 9 planes of 429 16-byte rows; TMA gathers each plane
//                 straight from XS in global memory with element strides (2, 2) -- no thread ever copies it.
// so DZ2 and DZ1 (82 KB per sample) never touch HBM and the weight gradients leave the SM once per launch.
// P1 is double-buffered and the TMA producer prefetches sample i+1's A2 / P1 / XS3 as soon as sample i releases the buffers.
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
constexpr int XS3_PLANE = 6912;               // 429 rows x 16 B, padded to 432 rows (rows 429..431 stay zero)
constexpr int SM_XS3 = 0;                     // 9 planes; the MMA's M = 128 also reads 7 junk "planes" out of the next buffer   62464
constexpr int SM_P1 = SM_XS3 + 62464;         // 2 x (448 rows x 128 B), SW128: P1, then DZ1 in place                          114688
constexpr int P1_BUF = 57344;
constexpr int SM_DZ2 = SM_P1 + 2 * P1_BUF;    // 16 zero halo rows + 512 rows, 64 B, SW64                                       33792
constexpr int SM_W2D = SM_DZ2 + 33792;        // [4 t][64 n][32 oc] bf16, SW64                                                  16384
constexpr int SM_TOTAL = SM_W2D + 16384;      // 227328
constexpr int P1_LOAD_ROWS = 448, A2_LOAD_ROWS = 432;   // 2 x 224, 2 x 216 row boxes
constexpr int K2_STEPS = 27;                  // 27 x 16 = 432 rows of the super-pixel row space

// TMEM columns
constexpr uint32_t TM_W2 = 0;                 // 64:  conv2.weight gradient  [ty][oc]
constexpr uint32_t TM_W1 = 64;                // 64:  conv1.weight gradient  lanes = patch value (72 used), columns = cell*16 + oc
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

__global__ void __launch_bounds__(FB_THREADS, 1) disc_bwd_fused_kernel(const __grid_constant__ CUtensorMap map_xs3, const __grid_constant__ CUtensorMap map_p1,
                                                                       const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_w,
                                                                       const FusedBwdArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_p1[2], full_a2, full_xs, dz2_ready, dz1_ready, mma1_done, mma2_done, wbar;
    __shared__ uint32_t tmem_s;
    __shared__ float red_s[48];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = a.B > (int)blockIdx.x ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;      // samples blockIdx.x, +gridDim.x, ...

    if (threadIdx.x == 0) {
        tc::mbar_init(&full_p1[0], 1); tc::mbar_init(&full_p1[1], 1); tc::mbar_init(&full_a2, 1); tc::mbar_init(&full_xs, 1); tc::mbar_init(&wbar, 1);
        tc::mbar_init(&dz2_ready, FB_WORKERS); tc::mbar_init(&dz1_ready, FB_WORKERS);
        tc::mbar_init(&mma1_done, 1); tc::mbar_init(&mma2_done, 1);
        tc::fence_barrier_init();
    }
    if (threadIdx.x < 48) red_s[threadIdx.x] = 0.f;
    if (warp == 1) { tc::tmem_alloc(&tmem_s, 512); tc::tmem_relinquish(); }
    {   // everything starts as zeros: DZ2's halo / tail rows and the planes' pad rows are never written and must read as 0
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < SM_W2D / 16; i += FB_THREADS) z[i] = make_uint4(0, 0, 0, 0);
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
            auto load_p1 = [&](int it) {                                           // sample `it` -> P1 buffer it & 1
                const int b = blockIdx.x + it * gridDim.x;
                unsigned char* dst = smem + SM_P1 + (it & 1) * P1_BUF;
                tc::mbar_expect_tx(&full_p1[it & 1], P1_LOAD_ROWS * 128);
                tc::tma_load_2d(dst, &map_p1, &full_p1[it & 1], 0, b * P1_ROWS);
                tc::tma_load_2d(dst + 224 * 128, &map_p1, &full_p1[it & 1], 0, b * P1_ROWS + 224);
            };
            load_p1(0);
            for (int it = 0; it < n_my; ++it) {
                const int b = blockIdx.x + it * gridDim.x;
                const uint32_t prev = (uint32_t)((it - 1) & 1);
                if (it > 0) tc::mbar_wait(&mma1_done, prev);                      // conv2 wgrad / dgrad MMAs of the previous sample have read DZ2
                tc::mbar_expect_tx(&full_a2, A2_LOAD_ROWS * 64);
                tc::tma_load_2d(smem + SM_DZ2 + 1024, &map_a2, &full_a2, 0, b * P1_ROWS);
                tc::tma_load_2d(smem + SM_DZ2 + 1024 + 216 * 64, &map_a2, &full_a2, 0, b * P1_ROWS + 216);
                if (it > 0) tc::mbar_wait(&mma2_done, prev);                      // its conv1 wgrad MMAs have read XS3 and the other P1 buffer (DZ1)
                tc::mbar_expect_tx(&full_xs, 9 * P1_ROWS * 16);
#pragma unroll
                for (int pl = 0; pl < 9; ++pl)                                    // plane (ay, ax): XS super pixel (2 sy - 1 + ay, 2 sx - 1 + ax), zero outside
                    tc::tma_load_4d(smem + SM_XS3 + pl * XS3_PLANE, &map_xs3, &full_xs, 0, pl % 3 - 1, pl / 3 - 1, b);
                if (it + 1 < n_my) load_p1(it + 1);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (n_my > 0 && tc::elect_one()) {
            const uint32_t leader = 1;
            constexpr uint64_t P1_MN = tc::smem_desc_base(128, 1024, tc::SW_128B);    // conv2 wgrad A: atom 1 = one row (128 B) later
            constexpr uint64_t DZ2_MN = tc::smem_desc_base(0, 512, tc::SW_64B);       // conv2 wgrad B
            constexpr uint64_t KM64 = tc::smem_desc_base(0, 512, tc::SW_64B);         // conv2 dgrad A (DZ2 rows) and B (W2d)
            constexpr uint64_t XS3_MN = tc::smem_desc_base(128, XS3_PLANE, tc::SW_NONE);   // conv1 wgrad A: 8-row K groups 128 B apart, M atoms one plane apart
            constexpr uint64_t DZ1_MN = tc::smem_desc_base(128, 1024, tc::SW_128B);        // conv1 wgrad B: DZ1 rows (64 values), one atom
            constexpr uint32_t ID_WG2 = tc::idesc_bf16(128, 32, 1, 1), ID_DG = tc::idesc_bf16(128, 64), ID_WG1 = tc::idesc_bf16(128, 64, 1, 1);
            const uint32_t p1_base = tc::smem_u32(smem + SM_P1), dz2 = tc::smem_u32(smem + SM_DZ2) + 1024, w2d = tc::smem_u32(smem + SM_W2D);
            const uint32_t xs3 = tc::smem_u32(smem + SM_XS3);
            tc::mbar_wait(&wbar, 0);
            for (int it = 0; it < n_my; ++it) {
                const uint32_t ph = (uint32_t)(it & 1);
                const uint32_t p1 = p1_base + (it & 1) * P1_BUF;
                tc::mbar_wait(&dz2_ready, ph);
                tc::mbar_wait(&full_p1[it & 1], (uint32_t)((it >> 1) & 1));
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
                if (!(a.dbg_skip & 4)) {
#pragma unroll 9
                    for (int k = 0; k < K2_STEPS; ++k)
                        tc::mma_f16_ss_pred(tmem + TM_W1, tc::smem_desc(XS3_MN, xs3 + k * 256), tc::smem_desc(DZ1_MN, p1 + k * 16 * 128), ID_WG1, (it | k) != 0, leader);
                }
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
        const uint32_t dz2s = tc::smem_u32(smem + SM_DZ2) + 1024, p1_base = tc::smem_u32(smem + SM_P1);
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
            // ---- W3: conv2 dgrad epilogue -> DZ1 in place over P1, conv1.bias gradient
            tc::mbar_wait(&mma1_done, ph);
            tc::tc_fence_after();
            const uint32_t p1s = p1_base + (it & 1) * P1_BUF;
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {
                const int R = tile * 128 + tl;
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem + tlane + TM_DG + tile * 64 + h * 32, r);      // cells (dy = h, dx = 0 | 1) x 16 channels
                tc::tmem_ld_wait();
                if (R >= A2_LOAD_ROWS) continue;
                const int sy = R / P1_W, sx = R - sy * P1_W;
                const int oy = 2 * sy + h - 1;
                const uint32_t prow = p1s + R * 128;
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int ox = 2 * sx + dx - 1;
                    // zero-padding cells of P1 have no conv1 output behind them; rows 429..431 hold the next sample's P1: both become 0
                    const bool cell = R < P1_ROWS && oy >= 0 && oy < 64 && ox >= 0 && ox < 25;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {                  // 8 channels per 16-byte chunk
                        const uint32_t addr = prow + ((((h * 2 + dx) * 2 + hh) ^ (R & 7)) << 4);
                        uint32_t o[4] = {0u, 0u, 0u, 0u};
                        if (cell) {
                            const uint4 av = tc::lds128(addr);
                            const uint32_t au[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = dx * 16 + hh * 8 + 2 * j;
                                const float g0 = __uint_as_float(r[c]) * (bf_lo(au[j]) > 0.f ? 1.f : 0.2f);
                                const float g1 = __uint_as_float(r[c + 1]) * (bf_hi(au[j]) > 0.f ? 1.f : 0.2f);
                                o[j] = pack_bf16x2(g0, g1);
                                db1[hh * 8 + 2 * j] += bf_lo(o[j]);
                                db1[hh * 8 + 2 * j + 1] += bf_hi(o[j]);
                            }
                        }
                        tc::sts128(addr, make_uint4(o[0], o[1], o[2], o[3]));
                    }
                }
            }
            tc::tc_fence_before();
            tc::fence_proxy_async_smem();
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
            }
            {   // conv1: TMEM lane m = patch value ((ay*3+ax)*8 + (dy',dx',ch)), column = cell*16 + oc; patch pixel (py,px) = (2ay+dy', 2ax+dx')
                // feeds cell (dy,dx) through tap (ky,kx) = (py - 2dy, px - 2dx)  ->  conv1.weight[oc][ch][ky][kx]
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem + tlane + TM_W1 + h * 32, r);                  // cells 2h, 2h+1
                tc::tmem_ld_wait();
                if (tl < 72) {
                    const int at = tl >> 3, e = tl & 7, py = 2 * (at / 3) + (e >> 2), px = 2 * (at % 3) + ((e >> 1) & 1), ch = e & 1;
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        const int ky = py - 2 * h, kx = px - 2 * dx;
                        if (ky < 0 || ky > 3 || kx < 0 || kx > 3) continue;
#pragma unroll
                        for (int oc = 0; oc < 16; ++oc) atomicAdd(&a.dw1[((oc * 2 + ch) * 4 + ky) * 4 + kx], __uint_as_float(r[dx * 16 + oc]));
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
    {   // XS as (B, 65, 26, 8): every second super pixel in both directions -> one 33 x 13 plane of 16-byte rows per load
        const uint64_t dims[4] = {8, (uint64_t)XS_W, 65, (uint64_t)B}, strides[3] = {16, 16 * XS_W, 16 * (uint64_t)XS_ROWS};
        const uint32_t box[4] = {8, 26, 65, 1}, estr[4] = {1, 2, 2, 1};
        MMG_REQUIRE(tc::make_map_nd_bf16(&map_xs, xs, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE) == 0, MMG_EINVAL, "disc_bwd_fused: tensor map (xs3)");
    }
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
