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
#ifdef MMG_ABLATION
#include <stdlib.h>      // timing experiments only: -DMMG_ABLATION builds read MMG_DBG_SKIP / MMG_DBG_SKIP_FWD
#endif

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
    __shared__ uint64_t full_p1[2], full_a2, full_xs, dz2_ready[4], dz1_ready, dg_done[4], mma2_done, wbar;    // [4]: one per 128-row tile
    __shared__ uint32_t tmem_s;
    __shared__ float red_s[48];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = a.B > (int)blockIdx.x ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;      // samples blockIdx.x, +gridDim.x, ...

    if (threadIdx.x == 0) {
        tc::mbar_init(&full_p1[0], 1); tc::mbar_init(&full_p1[1], 1); tc::mbar_init(&full_a2, 1); tc::mbar_init(&full_xs, 1); tc::mbar_init(&wbar, 1);
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&dz2_ready[i], FB_WORKERS); tc::mbar_init(&dg_done[i], 1); }
        tc::mbar_init(&dz1_ready, FB_WORKERS); tc::mbar_init(&mma2_done, 1);
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
                if (it > 0) tc::mbar_wait(&dg_done[3], prev);                     // conv2 wgrad / dgrad MMAs of the previous sample have read DZ2
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
                tc::mbar_wait(&full_p1[it & 1], (uint32_t)((it >> 1) & 1));
                // tile by tile behind the workers' DZ2 pass: as soon as rows [128 t, 128 t + 128) of DZ2 exist, their conv2 wgrad K steps and
                // dgrad tile t are issued, and the commit of tile t lets the dgrad epilogue of that tile start while later tiles still run
#pragma unroll
                for (int tile = 0; tile < 4; ++tile) {
                    tc::mbar_wait(&dz2_ready[tile], ph);
                    tc::tc_fence_after();
                    const int k_lo = tile * 8, k_hi = tile == 3 ? K2_STEPS : tile * 8 + 8;
                    if (!(a.dbg_skip & 1)) {
#pragma unroll
                        for (int ty = 0; ty < 2; ++ty)
#pragma unroll
                            for (int k = k_lo; k < k_hi; ++k)
                                tc::mma_f16_ss_pred(tmem + TM_W2 + ty * 32, tc::smem_desc(P1_MN, p1 + (ty * P1_W + k * 16) * 128),
                                                    tc::smem_desc(DZ2_MN, dz2 + k * 16 * 64), ID_WG2, (it | k) != 0, leader);
                    }
                    if (!(a.dbg_skip & 2)) {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                tc::mma_f16_ss_pred(tmem + TM_DG + tile * 64, tc::smem_desc(KM64, dz2 + (tile * 128 - ((t >> 1) * P1_W + (t & 1))) * 64 + k * 32),
                                                    tc::smem_desc(KM64, w2d + t * 4096 + k * 32), ID_DG, (t | k) != 0, leader);
                    }
                    tc::mma_commit_pred(&dg_done[tile], leader);
                }
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
                if (tile * 128 + q * 32 >= A2_LOAD_ROWS) {            // warp-uniform: this warp's 32 rows are all beyond the loaded rows
                    tc::mbar_arrive(&dz2_ready[tile]);
                    continue;
                }
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
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&dz2_ready[tile]);                    // the MMAs of this tile may start
            }
            tmem_st_wait();
            // ---- W3: conv2 dgrad epilogue -> DZ1 in place over P1, conv1.bias gradient (tile t as soon as its MMAs have committed)
            const uint32_t p1s = p1_base + (it & 1) * P1_BUF;
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {
                tc::mbar_wait(&dg_done[tile], ph);
                tc::tc_fence_after();
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
#ifdef MMG_ABLATION
    { const char* e = getenv("MMG_DBG_SKIP"); a.dbg_skip = e ? atoi(e) : 0; }
#else
    a.dbg_skip = 0;
#endif
    const int grid = (int)(B < MMG_NUM_SMS ? B : MMG_NUM_SMS);
    MMG_CUDA(cudaFuncSetAttribute(disc_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL + 1024));
    disc_bwd_fused_kernel<<<grid, FB_THREADS, SM_TOTAL + 1024, (cudaStream_t)stream>>>(map_xs, map_p1, map_a2, map_w, a);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"

// ================================================================================================================
// Fused forward of the bf16 tensor-core discriminator (DiscriminatorCNN.forward, network_tests.py:156-160): ONE persistent
// kernel per pass instead of xs_pack + conv1_fwd + conv2_fwd.  A CTA walks over whole samples:
//   S1  workers   X (u8 staged through shared memory by a bulk copy, or fp32 read directly) -> XS rows in shared memory (+ global, for the backward)
//   S2  tcgen05   conv1: 14 row tiles x 2 taps (M128 x N16 x K16) -> 224 TMEM columns
//   S3  workers   bias + LeakyReLU -> bf16 -> P1 super-pixel rows in shared memory (128-byte swizzle); one TMA store sends P1 to global
//   S4  tcgen05   conv2: 4 row tiles x 4 taps x 4 K steps (M128 x N32 x K16), tap-shifted descriptors over the SAME P1 rows
//   S5  workers   bias + LeakyReLU -> A2 rows in a shared-memory staging tile (one TMA store per sample); fc partial dots reduced per sample ->
//                 logits[b] (with the fc bias; no global atomics)
// P1 never makes the HBM round trip between conv1 and conv2.  HBM per sample: 12.8 KB read, 110 KB written (what the backward reads).
// ================================================================================================================
namespace {

constexpr int FF_XS_ROWS = 1824;                  // 14 tiles x 128 + 27 halo rows, padded
constexpr int FS_XS = 0;                          // 29184 -> 29696
constexpr int FS_P1 = 29696;                      // 528 rows x 128 B (rows >= 429 stay zero)   67584
constexpr int FS_W2 = FS_P1 + 67584;              // [4 t][32 oc][64 k] bf16 SW128              16384
constexpr int FS_W1 = FS_W2 + 16384;              // [2 ty][16 oc][16 k] bf16 SW32              1024
constexpr int FS_X = FS_W1 + 1024;                // staged u8 input                            12800
constexpr int FS_WF = FS_X + 12800;               // fc.weight slices, fp32, [8 (h, c4)][432 rows][4]: conflict-free LDS.128   55296
constexpr int FS_A2 = ((FS_WF + 8 * 432 * 16 + 1023) / 1024) * 1024;      // A2 staging, 432 rows x 64 B, SW64 (TMA-stored once per sample)   27648
constexpr int FS_TOTAL = FS_A2 + 432 * 64;        // 210944
constexpr uint32_t TF_C1 = 0, TF_C2 = 256;        // TMEM: conv1 14 x 16 columns, conv2 4 x 32 columns

struct FusedFwdArgs {
    const void* x; int x_f32; const int64_t* x_index;       // x_index != NULL: sample b is row x_index[b] of x (the HBM-resident training set)
    const float* b1; const float* b2; const float* wfcp; const float* bfc;
    __nv_bfloat16* xs; __nv_bfloat16* a2; float* logits;
    int B; int dbg_skip;      // dbg_skip (env MMG_DBG_SKIP_FWD, timing experiments only): bit0 conv1, bit1 conv2 MMAs are not issued
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(FB_THREADS, 1) disc_fwd_fused_kernel(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
                                                                       const __grid_constant__ CUtensorMap map_p1a, const __grid_constant__ CUtensorMap map_p1b,
                                                                       const __grid_constant__ CUtensorMap map_a2a, const __grid_constant__ CUtensorMap map_a2b,
                                                                       const FusedFwdArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t x_full, x_empty, xs_ready, c1_done, p1_ready[4], c2_done[4], wbar;     // [4]: one per 128-row conv2 tile
    __shared__ uint32_t tmem_s;
    __shared__ float logit_s;
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = a.B > (int)blockIdx.x ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        tc::mbar_init(&x_full, 1); tc::mbar_init(&wbar, 1); tc::mbar_init(&c1_done, 1);
        tc::mbar_init(&x_empty, FB_WORKERS); tc::mbar_init(&xs_ready, FB_WORKERS);
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&p1_ready[i], FB_WORKERS); tc::mbar_init(&c2_done[i], 1); }
        tc::fence_barrier_init();
        logit_s = 0.f;
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_s, 512); tc::tmem_relinquish(); }
    {   // XS halo rows and P1 pad cells / tail rows are never written: they must read as zeros
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < FS_W2 / 16; i += FB_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp == 0) {
        if (lane == 0 && n_my > 0) {
            tc::mbar_expect_tx(&wbar, 16384 + 1024);
            tc::tma_load_2d(smem + FS_W2, &map_w2, &wbar, 0, 0);
            tc::tma_load_2d(smem + FS_W1, &map_w1, &wbar, 0, 0);
            if (!a.x_f32) {
                for (int it = 0; it < n_my; ++it) {
                    const int b = blockIdx.x + it * gridDim.x;
                    if (it > 0) tc::mbar_wait(&x_empty, (uint32_t)((it - 1) & 1));
                    tc::mbar_expect_tx(&x_full, 12800);
                    bulk_load_1d(smem + FS_X, (const unsigned char*)a.x + (size_t)(a.x_index ? a.x_index[b] : b) * 12800, 12800, &x_full);
                }
            }
        }
    } else if (warp == 1) {
        if (n_my > 0 && tc::elect_one()) {
            constexpr uint64_t XS_K = tc::smem_desc_base(16, 128, tc::SW_NONE);      // conv1 A: K chunk 1 = the next 16-byte row
            constexpr uint64_t W1_K = tc::smem_desc_base(0, 256, tc::SW_32B);
            constexpr uint64_t KM128 = tc::smem_desc_base(0, 1024, tc::SW_128B);     // conv2 A (P1 rows) and B (W2p)
            constexpr uint32_t ID_C1 = tc::idesc_bf16(128, 16), ID_C2 = tc::idesc_bf16(128, 32);
            const uint32_t xs = tc::smem_u32(smem + FS_XS), p1 = tc::smem_u32(smem + FS_P1), w1 = tc::smem_u32(smem + FS_W1), w2 = tc::smem_u32(smem + FS_W2);
            tc::mbar_wait(&wbar, 0);
            for (int it = 0; it < n_my; ++it) {
                const uint32_t ph = (uint32_t)(it & 1);
                tc::mbar_wait(&xs_ready, ph);
                tc::tc_fence_after();
#pragma unroll
                for (int tile = 0; tile < 14; ++tile)
#pragma unroll
                    for (int ty = 0; ty < 2; ++ty)
                        if (!(a.dbg_skip & 1)) tc::mma_f16_ss(tmem + TF_C1 + tile * 16, tc::smem_desc(XS_K, xs + (tile * 128 + ty * XS_W) * 16), tc::smem_desc(W1_K, w1 + ty * 512), ID_C1, ty != 0);
                tc::mma_commit(&c1_done);
                // tile by tile behind the conv1 epilogue: conv2 tile t starts as soon as the P1 rows it reads (<= 128 t + 141) exist, and its
                // commit lets the conv2 epilogue of that tile run while the later tiles are still in the tensor pipe
#pragma unroll
                for (int tile = 0; tile < 4; ++tile) {
                    tc::mbar_wait(&p1_ready[tile], ph);
                    tc::tc_fence_after();
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (!(a.dbg_skip & 2)) tc::mma_f16_ss(tmem + TF_C2 + tile * 32, tc::smem_desc(KM128, p1 + (tile * 128 + (t >> 1) * P1_W + (t & 1)) * 128 + k * 32),
                                           tc::smem_desc(KM128, w2 + t * 4096 + k * 32), ID_C2, (t | k) != 0);
                    tc::mma_commit(&c2_done[tile]);
                }
            }
        }
    } else {
        const int q = warp & 3, h = (warp - 2) >> 2, tl = q * 32 + lane, w = threadIdx.x - 64;
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        const uint32_t xs_s = tc::smem_u32(smem + FS_XS), p1_s = tc::smem_u32(smem + FS_P1), x_s = tc::smem_u32(smem + FS_X), a2_s = tc::smem_u32(smem + FS_A2);
        float b1r[16], b2r[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { b1r[c] = a.b1[c]; b2r[c] = a.b2[h * 16 + c]; }
        const uint32_t wf_s = tc::smem_u32(smem + FS_WF);
        for (int idx = w; idx < 8 * 432; idx += FB_WORKERS) {         // fc.weight (junk rows carry 0) -> shared memory, once per CTA
            const int g = idx / 432, R = idx - g * 432;
            const float4 v = R < P1_ROWS ? reinterpret_cast<const float4*>(a.wfcp + R * 32)[g] : make_float4(0.f, 0.f, 0.f, 0.f);
            tc::sts128(wf_s + idx * 16, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float bfc = a.bfc[0];
        // ---- S1: XS rows (8 values = (dy,dx,ch) of super pixel (sy,sx) of the zero-padded input), to shared memory and to global
        // Sample 0's rows also go to global memory row by row; later samples leave the SM as ONE bulk copy of the finished 27 KB (issued by
        // thread 0 behind the barrier that ends the sample's conv2 epilogue), which keeps 1690 16-byte stores out of the workers' instruction stream.
        auto build_xs = [&](int it) {
            const int b = blockIdx.x + it * gridDim.x;
            if (!a.x_f32) tc::mbar_wait(&x_full, (uint32_t)(it & 1));
            if (!a.x_f32) {
                // all of this thread's byte loads (7 rows x 8) are issued before any is used: one shared-memory latency instead of seven
                constexpr int NR = (XS_ROWS + FB_WORKERS - 1) / FB_WORKERS;
                uint32_t u[NR][8];
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int rr = w + i * FB_WORKERS;
                    const int sy = rr / XS_W, sx = rr - sy * XS_W;
                    const int iy0 = 2 * sy - 1, ix0 = 2 * sx - 1, off0 = iy0 * 50 + ix0;      // element (dy,dx,ch) sits at off0 + dy*50 + dx + ch*6400
                    const bool row = rr < XS_ROWS, y0 = iy0 >= 0, y1 = iy0 + 1 < 128, x0 = ix0 >= 0, x1 = ix0 + 1 < 50;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int dy = e >> 2, dx = (e >> 1) & 1, ch = e & 1;
                        const bool in = row && (dy ? y1 : y0) && (dx ? x1 : x0);
                        u[i][e] = in ? tc::lds_u8(x_s + off0 + dy * 50 + dx + ch * 6400) : 0u;
                    }
                }
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int rr = w + i * FB_WORKERS;
                    if (rr >= XS_ROWS) break;
                    uint32_t f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __float_as_uint((float)u[i][e]);      // integers 0..255 are exact in bf16 = the high half of the float
                    const uint4 pk = make_uint4(__byte_perm(f[0], f[1], 0x7632), __byte_perm(f[2], f[3], 0x7632), __byte_perm(f[4], f[5], 0x7632), __byte_perm(f[6], f[7], 0x7632));
                    tc::sts128(xs_s + rr * 16, pk);
                    if (it == 0) *reinterpret_cast<uint4*>(a.xs + ((size_t)b * XS_ROWS + rr) * 8) = pk;
                }
            } else {
                const long long xb = a.x_index ? a.x_index[b] : b;
                for (int rr = w; rr < XS_ROWS; rr += FB_WORKERS) {
                    const int sy = rr / XS_W, sx = rr - sy * XS_W;
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int iy = 2 * sy + (e >> 2) - 1, ix = 2 * sx + ((e >> 1) & 1) - 1, ch = e & 1;
                        v[e] = (iy >= 0 && iy < 128 && ix >= 0 && ix < 50) ? reinterpret_cast<const float*>(a.x)[(size_t)xb * 12800 + (ch * 128 + iy) * 50 + ix] : 0.f;
                    }
                    const uint4 pk = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    tc::sts128(xs_s + rr * 16, pk);
                    if (it == 0) *reinterpret_cast<uint4*>(a.xs + ((size_t)b * XS_ROWS + rr) * 8) = pk;
                }
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&xs_ready);
            tc::mbar_arrive(&x_empty);
        };
        if (n_my > 0) build_xs(0);
        for (int it = 0; it < n_my; ++it) {
            const int b = blockIdx.x + it * gridDim.x;
            const uint32_t ph = (uint32_t)(it & 1);
            // ---- S3: conv1 epilogue -> P1 (rows = super pixels, 64 values = (dy,dx,c16)); even tiles for h = 0, odd tiles for h = 1
            tc::mbar_wait(&c1_done, ph);
            tc::tc_fence_after();
            if (w == 0) tc::bulk_wait_group_read<1>();                // the previous sample's P1 store (every group but the newest) has finished reading shared memory
            asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
            for (int tt = 0; tt < 7; ++tt) {
                const int tile = 2 * tt + h;
                uint32_t r[16];
                tc::tmem_ld_32x16(tmem + tlane + TF_C1 + tile * 16, r);
                tc::tmem_ld_wait();
                const int m = tile * 128 + tl, oy = m / XS_W, ox = m - oy * XS_W;
                if (m < XS_ROWS && oy < 64 && ox < 25) {
                    uint32_t o[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float z0 = __uint_as_float(r[2 * c]) + b1r[2 * c], z1 = __uint_as_float(r[2 * c + 1]) + b1r[2 * c + 1];
                        o[c] = pack_bf16x2(fmaxf(z0, 0.2f * z0), fmaxf(z1, 0.2f * z1));       // LeakyReLU(0.2) = max(z, 0.2 z)
                    }
                    const int yp = oy + 1, xp = ox + 1, R = (yp >> 1) * P1_W + (xp >> 1), cell = (yp & 1) * 2 + (xp & 1);
                    const uint32_t rowp = p1_s + R * 128;
                    tc::sts128(rowp + (((2 * cell) ^ (R & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
                    tc::sts128(rowp + (((2 * cell + 1) ^ (R & 7)) << 4), make_uint4(o[4], o[5], o[6], o[7]));
                }
                // conv1 tiles <= 2 tt + 1 are in P1 once every worker is past this point: conv2 tile 0 reads P1 rows <= 141 (conv1 rows <= 544:
                // tiles 0..4), tile 1 rows <= 269 (tiles 0..8), tiles 2 and 3 everything
                if (tt == 2 || tt == 4 || tt == 6) {
                    tc::tc_fence_before();
                    tc::fence_proxy_async_smem();
                    tc::mbar_arrive(&p1_ready[tt == 2 ? 0 : tt == 4 ? 1 : 2]);
                    if (tt == 6) tc::mbar_arrive(&p1_ready[3]);
                }
            }
            if (w == 0) tc::bulk_wait_group_read<0>();                // the previous sample's A2 store and this sample's XS store (issued one epilogue ago) are done
            asm volatile("bar.sync 1, 256;" ::: "memory");            // with shared memory: the A2 staging rows and the XS rows may be overwritten below
            if (w == 0) {                                             // P1 -> global for the backward (rows b*429 .. +429, two boxes)
                tc::tma_store_2d(&map_p1a, smem + FS_P1, 0, b * P1_ROWS);
                tc::tma_store_2d(&map_p1b, smem + FS_P1 + 224 * 128, 0, b * P1_ROWS + 224);
                tc::bulk_commit_group();
            }
            if (it + 1 < n_my) build_xs(it + 1);                      // conv1 of this sample is done with XS: the next sample's rows are built under conv2's MMAs
            // ---- S5: conv2 epilogue -> A2 rows (global), fc partial dot
            float dot = 0.f;
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {
                tc::mbar_wait(&c2_done[tile], ph);
                tc::tc_fence_after();
                const int R = tile * 128 + tl;
                uint32_t r[16];
                tc::tmem_ld_32x16(tmem + tlane + TF_C2 + tile * 32 + h * 16, r);
                tc::tmem_ld_wait();
                if (R >= P1_ROWS) continue;
                const int oy = R / P1_W, ox = R - oy * P1_W;
                const bool real = oy < 32 && ox < 12;                 // junk rows of the row space are stored as zeros
                uint32_t o[8];
                float wv[16];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const uint4 q4 = tc::lds128(wf_s + ((h * 4 + c4) * 432 + R) * 16);
                    wv[4 * c4] = __uint_as_float(q4.x); wv[4 * c4 + 1] = __uint_as_float(q4.y); wv[4 * c4 + 2] = __uint_as_float(q4.z); wv[4 * c4 + 3] = __uint_as_float(q4.w);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float z0 = __uint_as_float(r[2 * c]) + b2r[2 * c], z1 = __uint_as_float(r[2 * c + 1]) + b2r[2 * c + 1];
                    z0 = fmaxf(z0, 0.2f * z0); z1 = fmaxf(z1, 0.2f * z1);
                    o[c] = real ? pack_bf16x2(z0, z1) : 0u;
                    dot = fmaf(bf_lo(o[c]), wv[2 * c], dot);                              // the bf16 values the backward will read
                    dot = fmaf(bf_hi(o[c]), wv[2 * c + 1], dot);
                }
                const uint32_t rowp = a2_s + R * 64;                  // A2 row -> staging (64-byte swizzle), stored by TMA below: full lines instead of half-sector stores
                const int sw = (R >> 1) & 3;
                tc::sts128(rowp + (((2 * h) ^ sw) << 4), make_uint4(o[0], o[1], o[2], o[3]));
                tc::sts128(rowp + (((2 * h + 1) ^ sw) << 4), make_uint4(o[4], o[5], o[6], o[7]));
            }
            dot = warp_sum(dot);
            if (lane == 0) atomicAdd(&logit_s, dot);
            tc::tc_fence_before();
            tc::fence_proxy_async_smem();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (w == 0) {
                a.logits[b] = logit_s + bfc; logit_s = 0.f;
                tc::tma_store_2d(&map_a2a, smem + FS_A2, 0, b * P1_ROWS);                 // A2 rows b*429 .. +429 (two boxes)
                tc::tma_store_2d(&map_a2b, smem + FS_A2 + 216 * 64, 0, b * P1_ROWS + 216);
                if (it + 1 < n_my)                                    // the next sample's XS rows (built above by all workers)
                    tc::bulk_store_1d(a.xs + (size_t)(b + gridDim.x) * XS_ROWS * 8, xs_s, XS_ROWS * 16);
                tc::bulk_commit_group();
            }
            // (the next use of logit_s comes after the next sample's two bar.sync: ordered)
        }
        if (w == 0) tc::bulk_wait_group_read<0>();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" {

// x (B,2,128,50) uint8 (x_dtype 2) or float32 (x_dtype 0) -> xs (B*1690,8), p1 (B*429,64), a2 (B*429,32) bf16 and logits (B,) fp32
// (fc bias included; nothing needs to be pre-initialised).
int mmg_disc_fwd_fused_gather(const void* x, int x_dtype, const int64_t* x_index, const void* packed, const float* conv1_b, const float* conv2_b,
                              const float* fc_b, void* xs, void* p1, void* a2, float* logits, int64_t B, void* stream);

int mmg_disc_fwd_fused(const void* x, int x_dtype, const void* packed, const float* conv1_b, const float* conv2_b, const float* fc_b, void* xs, void* p1,
                       void* a2, float* logits, int64_t B, void* stream) {
    return mmg_disc_fwd_fused_gather(x, x_dtype, nullptr, packed, conv1_b, conv2_b, fc_b, xs, p1, a2, logits, B, stream);
}

// The same with a gather: sample b of the pass is row x_index[b] (int64, device memory) of x -- the discriminator reads the real rolls
// straight out of the HBM-resident training set by sampler index, no gathered copy of the batch is ever made.  x_index == NULL: rows 0..B-1.
int mmg_disc_fwd_fused_gather(const void* x, int x_dtype, const int64_t* x_index, const void* packed, const float* conv1_b, const float* conv2_b,
                              const float* fc_b, void* xs, void* p1, void* a2, float* logits, int64_t B, void* stream) {
    MMG_REQUIRE(x && packed && conv1_b && conv2_b && fc_b && xs && p1 && a2 && logits && B >= 0, MMG_EINVAL, "disc_fwd_fused: bad arguments");
    MMG_REQUIRE(x_dtype == 0 || x_dtype == 2, MMG_EINVAL, "disc_fwd_fused: x_dtype must be 0 (f32) or 2 (u8)");
    if (B == 0) return MMG_OK;
    MMG_REQUIRE(B * XS_ROWS < (1LL << 31) - 4096, MMG_EUNSUPPORTED, "disc_fwd_fused: batch too large");
    MMG_REQUIRE(x_dtype != 2 || ((uintptr_t)x & 15) == 0, MMG_EINVAL, "disc_fwd_fused: x must be 16-byte aligned");
    const unsigned char* pk = (const unsigned char*)packed;
    CUtensorMap map_w1, map_w2, map_p1a, map_p1b, map_a2a, map_a2b;
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w1, pk, 16, 32, 32, 16, 32, CU_TENSOR_MAP_SWIZZLE_32B) == 0, MMG_EINVAL, "disc_fwd_fused: tensor map (w1b)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w2, pk + 2048, 64, 128, 128, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B) == 0, MMG_EINVAL, "disc_fwd_fused: tensor map (w2p)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_p1a, p1, 64, (uint64_t)(B * P1_ROWS), 128, 64, 224, CU_TENSOR_MAP_SWIZZLE_128B) == 0, MMG_EINVAL, "disc_fwd_fused: tensor map (p1 a)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_p1b, p1, 64, (uint64_t)(B * P1_ROWS), 128, 64, 205, CU_TENSOR_MAP_SWIZZLE_128B) == 0, MMG_EINVAL, "disc_fwd_fused: tensor map (p1 b)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_a2a, a2, 32, (uint64_t)(B * P1_ROWS), 64, 32, 216, CU_TENSOR_MAP_SWIZZLE_64B) == 0, MMG_EINVAL, "disc_fwd_fused: tensor map (a2 a)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_a2b, a2, 32, (uint64_t)(B * P1_ROWS), 64, 32, 213, CU_TENSOR_MAP_SWIZZLE_64B) == 0, MMG_EINVAL, "disc_fwd_fused: tensor map (a2 b)");
    FusedFwdArgs a;
    a.x = x; a.x_f32 = x_dtype == 0; a.x_index = x_index; a.b1 = conv1_b; a.b2 = conv2_b; a.wfcp = (const float*)(pk + 2048 + 32768); a.bfc = fc_b;
    a.xs = (__nv_bfloat16*)xs; a.a2 = (__nv_bfloat16*)a2; a.logits = logits; a.B = (int)B;
#ifdef MMG_ABLATION
    { const char* e = getenv("MMG_DBG_SKIP_FWD"); a.dbg_skip = e ? atoi(e) : 0; }
#else
    a.dbg_skip = 0;
#endif
    const int grid = (int)(B < MMG_NUM_SMS ? B : MMG_NUM_SMS);
    MMG_CUDA(cudaFuncSetAttribute(disc_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_TOTAL + 1024));
    disc_fwd_fused_kernel<<<grid, FB_THREADS, FS_TOTAL + 1024, (cudaStream_t)stream>>>(map_w1, map_w2, map_p1a, map_p1b, map_a2a, map_a2b, a);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
