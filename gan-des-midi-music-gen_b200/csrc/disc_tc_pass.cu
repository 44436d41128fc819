// One whole discriminator pass of the MM-GAN loop body in ONE persistent tcgen05 kernel:
//   DiscriminatorCNN.forward (network_tests.py:156-160) -> BCEWithLogitsLoss (network_tests.py:248,304-306,313) -> autograd backward
//   (network_tests.py:307,314), per sample, with nothing but the uint8 roll read from HBM.
// BCE is per sample: dlogit[b] = (sigmoid(logit[b]) - y) / B is known the moment logit[b] is, so the activations the backward needs
// (XS 27 KB, P1 55 KB, A2 27 KB per sample; 110 KB written + 110 KB read per sample and pass by the two-kernel version in
// disc_tc_fused.cu) never leave the SM.  Layouts and descriptors are those of disc_tc.cu / disc_tc_fused.cu (validated on hardware).
//
// A CTA walks over whole samples; warp 0 = producer (bulk copies / TMA), warp 1 = one elected thread issuing tcgen05.mma, 8 worker warps.
// Per sample (W = workers, T = tensor pipe; every T phase is tile-pipelined behind the W phase that feeds it):
//   W  XSb   X (u8, bulk-copied to shared memory one sample ahead) -> XS rows (65 x 26 super pixels x 8 values, bf16)
//   T  C1    conv1: 14 row tiles x 2 MMAs (M128 N16 K16)                                   -> TMEM C1 (224 columns)
//   W  S3    bias + LeakyReLU -> bf16 -> P1 super-pixel rows (33 x 13 x 64 values, SW128)
//   T  C2    conv2: 4 tiles x 16 MMAs (M128 N32 K16) over tap-shifted descriptors of P1   -> TMEM C2 (aliases consumed C1 columns)
//   W  S5    bias + LeakyReLU -> bf16 -> A2 rows (SW64) + fc partial dots; CTA reduction -> logit -> BCE loss, dlogit
//   W  W1    A2 -> DZ2 in place (dlogit * fc.w * lrelu'), fc.weight gradient read-modify-written in TMEM, conv2.bias gradient
//   T  M1    conv2 weight gradient (P1 x DZ2, persistent TMEM accumulator) + conv2 data gradient (DZ2 tap-shift x W2d -> TMEM DG, aliases C1/C2)
//   W  W3    DG x lrelu'(P1) -> DZ1 in place over P1; conv1.bias gradient;  then XSb of the NEXT sample (under M2)
//   T  M2    conv1 weight gradient in the super-pixel row space: XS3 (3x3 neighbourhoods of XS super pixels, 9 planes) x DZ1
// XS3 is a strided gather of XS that only TMA can do for free: the producer parks XS in a per-CTA scratch slot (2 x 27 KB per CTA, 8 MB
// in total: it lives in L2, it is overwritten long before it would be evicted) and TMA-gathers the nine planes back.
// HBM per sample and pass: the 12.8 KB roll.  Weight gradients leave the SM once per launch.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int P1_ROWS = 429, P1_W = 13;       // 33 x 13 super pixels per sample (conv1 activations / conv2 row space)
constexpr int XS_ROWS = 1690, XS_W = 26;      // 65 x 26 super pixels per sample (input / conv1 row space)
constexpr int NWORK = 512;                    // warps 0-15: workers (4 warpgroups), warp 16 producer, warp 17 MMA, warps 18-19 idle (setmaxnreg needs whole warpgroups)
constexpr int NTHREADS = NWORK + 128;
constexpr int PRODUCER_WARP = 16, MMA_WARP = 17;

// shared-memory map (offsets from a 1024-byte aligned base)
constexpr int XS3_PLANE = 6912;               // 429 rows x 16 B, padded to 432 rows (rows 429..431 stay zero)
constexpr int SM_XS3 = 0;                     // 9 gathered planes + plane 9 = all ones (its 8 TMEM lanes collect the column sums of DZ1 = conv1.bias gradient)   69120
constexpr int SM_ONESA = SM_XS3 + 10 * XS3_PLANE;   // A tile of the bias MMAs: 8 rows [1, 1, 0 x 6] + 128 B of zeros (K chunk 1)                                256 (+256 pad)
constexpr int SM_P1 = 69632;                  // 448 rows x 128 B, SW128: P1, then DZ1 in place                                  57344
constexpr int SM_DZ2 = SM_P1 + 57344;         // 16 zero halo rows + 432 rows x 64 B, SW64: A2, then DZ2 in place                28672
constexpr int SM_W2 = SM_DZ2 + 28672;         // [4 t][32 oc][64 k] bf16 SW128 (conv2 forward B)                                 16384
constexpr int SM_W2D = SM_W2 + 16384;         // [4 t][64 n][32 oc] bf16 SW64 (conv2 dgrad B)                                    16384
constexpr int SM_W1 = SM_W2D + 16384;         // [2 ty][16 oc][16 k] bf16 SW32 (conv1 forward B)                                 1024
constexpr int SM_BB = SM_W1 + 1024;           // B tiles of the bias MMAs: conv1 [16 n][16 k] 512 B, conv2 [32 n][16 k] 1024 B: k0 = bias hi, k1 = bias lo (bf16), rest 0
constexpr int SM_XS = SM_BB + 1536;           // 1690 rows x 16 B, no swizzle (the junk rows of conv1's last tile read on into SM_X)  27040
constexpr int SM_X = SM_XS + 27040;           // staged uint8 roll                                                               12800
constexpr int SM_TOTAL = SM_X + 12800;        // 230816
constexpr int A2_ROWS = 432;                  // rows of the A2 / DZ2 tile that exist (27 K steps of 16)
constexpr int K2_STEPS = 27;
static_assert(SM_ONESA + 256 <= SM_P1 && SM_P1 % 1024 == 0 && SM_DZ2 % 1024 == 0 && SM_W2 % 1024 == 0 && SM_W2D % 1024 == 0 && SM_W1 % 256 == 0 && SM_BB % 128 == 0 &&
              SM_XS % 128 == 0 && SM_X % 16 == 0 && SM_TOTAL + 1024 + 608 <= 232448, "shared-memory map");

// TMEM columns
constexpr uint32_t TM_W2 = 0;                 // 64:  conv2.weight gradient  [ty][oc]            (persistent)
constexpr uint32_t TM_W1 = 64;                // 64:  conv1.weight gradient  lanes = patch value, lanes 72..79 = column sums of DZ1 (persistent)
constexpr uint32_t TM_FC = 128;               // 128: fc.weight gradient     [tile][oc]          (persistent)
constexpr uint32_t TM_X = 256;                // 256 transient columns:
                                              //   conv1 accumulators, tile t at TM_X + 16 t (14 tiles).  Tiles 0..7 of sample j + 1 are issued while sample j's
                                              //   backward still runs (its dgrad ring is in the upper half), tiles 8..13 once that ring has been read;
                                              //   conv2 accumulators, tile t at TM_X + 32 t (over conv1 tiles 0..7 once the conv1 epilogue has consumed them);
constexpr uint32_t TM_DGR = 384;              //   conv2 dgrad ring, 2 slots of 64 columns: tile t in slot t & 1 (over conv1 tiles 8..13, consumed long before)

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
// LeakyReLU'(a) for the two bf16 halves of a packed word, tested on the bits (a > 0 exactly as the float compare would)
__device__ __forceinline__ bool pos_lo(uint32_t u) { return (int)(u << 16) > 0; }
__device__ __forceinline__ bool pos_hi(uint32_t u) { return (int)u >= 0x10000; }

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

struct PassArgs {
    const void* x; int x_f32; const int64_t* x_index; long long x_rows;      // x_rows > 0: gathered row numbers are checked against it
    int* oob;                           // set to 1 when an index was out of range (the row is then read as row 0)
    const float* b1; const float* b2; const float* wfcp; const float* bfc;
    float y, inv_n;                     // BCE target of the pass; 1 / (rows behind the loss mean)
    float* logits; float* loss;         // optional outputs: logits (B,), loss[0] += sum_b bce_b * inv_n
    float* dw1; float* db1; float* dw2; float* db2; float* dwfc; float* dbfc;
    __nv_bfloat16* scratch;             // gridDim.x x 2 slots x 1690 rows x 8
    int B;
    volatile int* dbg;                  // progress words (host-mapped memory; debugging hangs), or NULL
};

#define PASS_DBG(slot, val) do { if (a.dbg) a.dbg[(blockIdx.x * 4 + (slot))] = (val); } while (0)
// timeline of CTA 0, samples 6 and 7 (debug entry point only): SM clock at the phase boundaries of worker thread 0 (k < 16) and of the MMA thread (16 + k)
#define PASS_TS(k) do { if (a.dbg && blockIdx.x == 0 && (it == 6 || it == 7)) a.dbg[1024 + (it - 6) * 32 + (k)] = (int)clock64(); } while (0)

// BIAS_MMA: the two convolution biases enter the accumulators through one extra MMA per tile (A = ones, B = bias hi | lo) instead of an
// FADD per value in the epilogues.
template <bool BIAS_MMA>
__global__ void __launch_bounds__(NTHREADS, 1) disc_pass_fused_kernel(const __grid_constant__ CUtensorMap map_xs3, const __grid_constant__ CUtensorMap map_w1,
                                                                      const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_w2d,
                                                                      const PassArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t wbar, x_full, xs_ready, xs_saved, full_xs3, c1a_done, c1b_done, p1_ready[4], c2_done[4], dz2_ready[4], dg_done[4], dg_read[2],
        dz1_ready[4], mma2_done, c2_read, wg1a_done;
    __shared__ uint32_t tmem_s;
    __shared__ __align__(16) float logit_part[2][16];     // per-warp partial fc dots of the current sample (two phases)
    __shared__ float red_s[48];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = a.B > (int)blockIdx.x ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;      // samples blockIdx.x, +gridDim.x, ...

    if (threadIdx.x == 0) {
        tc::mbar_init(&wbar, 1); tc::mbar_init(&x_full, 1); tc::mbar_init(&xs_saved, 1); tc::mbar_init(&full_xs3, 1);
        tc::mbar_init(&c1a_done, 1); tc::mbar_init(&c1b_done, 1); tc::mbar_init(&mma2_done, 1);
        tc::mbar_init(&xs_ready, NWORK); tc::mbar_init(&c2_read, NWORK); tc::mbar_init(&wg1a_done, 1);
        for (int i = 0; i < 4; ++i) {
            tc::mbar_init(&p1_ready[i], NWORK); tc::mbar_init(&c2_done[i], 1); tc::mbar_init(&dz2_ready[i], NWORK); tc::mbar_init(&dg_done[i], 1);
            tc::mbar_init(&dz1_ready[i], NWORK);
        }
        for (int i = 0; i < 2; ++i) tc::mbar_init(&dg_read[i], NWORK);
        tc::fence_barrier_init();
    }
    float* bias_s = reinterpret_cast<float*>(smem + SM_BB);      // !BIAS_MMA: conv1.bias (16) | conv2.bias (32) as fp32 where the bias tiles would be
    if (threadIdx.x < 48) red_s[threadIdx.x] = 0.f;
    if (warp == MMA_WARP) { tc::tmem_alloc(&tmem_s, 512); tc::tmem_relinquish(); }
    {   // everything that is never written must read as zeros (plane pad rows, P1 pad cells / tail rows, the DZ2 halo); constants: the ones plane,
        // the bias-MMA operands.  Every 16-byte word has exactly one writer here.
        for (int i = threadIdx.x; i < SM_W2 / 16; i += NTHREADS) {
            const int off = i * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (off >= SM_XS3 + 9 * XS3_PLANE && off < SM_ONESA) v = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);      // bf16 1.0 x 8
            else if (off >= SM_ONESA && off < SM_ONESA + 128) v = make_uint4(0x3F803F80u, 0, 0, 0);                                    // [1, 1, 0 x 6]
            *reinterpret_cast<uint4*>(smem + off) = v;
        }
        for (int i = threadIdx.x; i < (SM_X - SM_XS) / 16; i += NTHREADS) *reinterpret_cast<uint4*>(smem + SM_XS + i * 16) = make_uint4(0, 0, 0, 0);
        if (!BIAS_MMA) {
            if (threadIdx.x < 48) bias_s[threadIdx.x] = threadIdx.x < 16 ? a.b1[threadIdx.x] : a.b2[threadIdx.x - 16];
        } else
        for (int i = threadIdx.x; i < 1536 / 16; i += NTHREADS) {      // bias B tiles, K-major, no swizzle: core matrix (n group, k chunk) at n_g * 256 + k_c * 128
            const bool c2 = i >= 32;
            const int j = c2 ? i - 32 : i, n_g = j >> 4, k_c = (j >> 3) & 1, n = n_g * 8 + (j & 7);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (k_c == 0) {
                const float b = c2 ? a.b2[n] : a.b1[n];
                const __nv_bfloat16 hi = __float2bfloat16(b), lo = __float2bfloat16(b - __bfloat162float(hi));
                v.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
            }
            *reinterpret_cast<uint4*>(smem + SM_BB + i * 16) = v;
        }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp >= 16) {
        // (no setmaxnreg: the MMA issue loops must stay fully unrolled -- with run-time tile indices ptxas builds the descriptors in vector registers and
        // wraps every UTCHMMA in an ELECT / R2UR uniformisation loop, 40-70 cycles per MMA -- and unrolled they need ~88 registers; the pool of the CTA
        // is 640 x 96, so nothing is left to hand to the workers)
        if (warp == PRODUCER_WARP) {
            // ------------------------------------------------------------------ producer
            if (lane == 0 && n_my > 0) {
                tc::mbar_expect_tx(&wbar, 16384 + 16384 + 1024);
                tc::tma_load_2d(smem + SM_W2, &map_w2, &wbar, 0, 0);
                tc::tma_load_2d(smem + SM_W2D, &map_w2d, &wbar, 0, 0);
                tc::tma_load_2d(smem + SM_W1, &map_w1, &wbar, 0, 0);
                auto load_x = [&](int it) {
                    const int b = blockIdx.x + it * gridDim.x;
                    tc::mbar_expect_tx(&x_full, 12800);
                    long long row = a.x_index ? a.x_index[b] : b;
                    if (a.x_rows > 0 && (unsigned long long)row >= (unsigned long long)a.x_rows) { *a.oob = 1; row = 0; }
                    bulk_load_1d(smem + SM_X, (const unsigned char*)a.x + (size_t)row * 12800, 12800, &x_full);
                };
                if (!a.x_f32) load_x(0);
                for (int it = 0; it < n_my; ++it) {
                    const int slot = (int)blockIdx.x * 2 + (it & 1);
                    tc::mbar_wait(&xs_ready, (uint32_t)(it & 1));                     // XS(it) is complete; the staged roll has been consumed
                    PASS_DBG(0, it * 16 + 1);
                    if (!a.x_f32 && it + 1 < n_my) load_x(it + 1);
                    tc::bulk_store_1d(a.scratch + (size_t)slot * XS_ROWS * 8, tc::smem_u32(smem + SM_XS), XS_ROWS * 16);
                    tc::bulk_commit_group();
                    bulk_wait_group_all();                                            // the rows are in L2 (and shared memory has been read)
                    fence_proxy_async_all();
                    tc::mbar_arrive(&xs_saved);
                    PASS_DBG(0, it * 16 + 2);
                    if (it > 0) tc::mbar_wait(&mma2_done, (uint32_t)((it - 1) & 1));  // the conv1 wgrad MMAs of the previous sample have read XS3
                    tc::mbar_expect_tx(&full_xs3, 9 * P1_ROWS * 16);
#pragma unroll
                    for (int pl = 0; pl < 9; ++pl)                                    // plane (ay, ax): XS super pixel (2 sy - 1 + ay, 2 sx - 1 + ax), zero outside
                        tc::tma_load_4d(smem + SM_XS3 + pl * XS3_PLANE, &map_xs3, &full_xs3, 0, pl % 3 - 1, pl / 3 - 1, slot);
                    PASS_DBG(0, it * 16 + 3);
                }
            }
        } else if (warp == MMA_WARP) {
            // ------------------------------------------------------------------ MMA issuer
            if (n_my > 0 && tc::elect_one()) {
                const uint32_t leader = 1;
                // forward
                constexpr uint64_t XS_K = tc::smem_desc_base(16, 128, tc::SW_NONE);       // conv1 A: K chunk 1 = the next 16-byte row
                constexpr uint64_t W1_K = tc::smem_desc_base(0, 256, tc::SW_32B);
                constexpr uint64_t KM128 = tc::smem_desc_base(0, 1024, tc::SW_128B);      // conv2 A (P1 rows) and B (W2p)
                constexpr uint64_t ONES_K = tc::smem_desc_base(128, 0, tc::SW_NONE);      // bias A: every 8-row group is the same 128 bytes; K chunk 1 = zeros
                constexpr uint64_t BB_K = tc::smem_desc_base(128, 256, tc::SW_NONE);      // bias B
                constexpr uint32_t ID_C1 = tc::idesc_bf16(128, 16), ID_C2 = tc::idesc_bf16(128, 32);
                // backward
                constexpr uint64_t P1_MN = tc::smem_desc_base(128, 1024, tc::SW_128B);    // conv2 wgrad A: atom 1 = one row (128 B) later
                constexpr uint64_t DZ2_MN2 = tc::smem_desc_base(P1_W * 64, 512, tc::SW_64B);   // conv2 wgrad B: N atom 1 = thirteen rows (one super-pixel row) later
                constexpr uint64_t KM64 = tc::smem_desc_base(0, 512, tc::SW_64B);         // conv2 dgrad A (DZ2 rows) and B (W2d)
                constexpr uint64_t XS3_MN = tc::smem_desc_base(128, XS3_PLANE, tc::SW_NONE);   // conv1 wgrad A: 8-row K groups 128 B apart, M atoms one plane apart
                constexpr uint64_t DZ1_MN = tc::smem_desc_base(128, 1024, tc::SW_128B);        // conv1 wgrad B: DZ1 rows (64 values), one atom
                constexpr uint32_t ID_WG2 = tc::idesc_bf16(128, 64, 1, 1), ID_DG = tc::idesc_bf16(128, 64), ID_WG1 = tc::idesc_bf16(128, 64, 1, 1);
                const uint32_t xs = tc::smem_u32(smem + SM_XS), p1 = tc::smem_u32(smem + SM_P1), w1 = tc::smem_u32(smem + SM_W1), w2 = tc::smem_u32(smem + SM_W2);
                const uint32_t dz2 = tc::smem_u32(smem + SM_DZ2) + 1024, w2d = tc::smem_u32(smem + SM_W2D), xs3 = tc::smem_u32(smem + SM_XS3);
                const uint32_t onesa = tc::smem_u32(smem + SM_ONESA), bb1 = tc::smem_u32(smem + SM_BB), bb2 = bb1 + 512;
                tc::mbar_wait(&wbar, 0);
                // conv1 tile `tile` of the sample whose XS rows are in shared memory
                auto conv1_tile = [&](int tile) {
                    if (BIAS_MMA) tc::mma_f16_ss_pred(tmem + TM_X + tile * 16, tc::smem_desc(ONES_K, onesa), tc::smem_desc(BB_K, bb1), ID_C1, 0, leader);
#pragma unroll
                    for (int ty = 0; ty < 2; ++ty)
                        tc::mma_f16_ss_pred(tmem + TM_X + tile * 16, tc::smem_desc(XS_K, xs + (tile * 128 + ty * XS_W) * 16), tc::smem_desc(W1_K, w1 + ty * 512), ID_C1,
                                            (BIAS_MMA || ty != 0) ? 1u : 0u, leader);
                };
                // conv1 wgrad K steps of P1-row tile `part` (+ column sums of DZ1 in the lanes of the ones plane)
                auto wgrad1_part = [&](int it, int part) {
                    const int k_lo = part * 8, k_hi = part == 3 ? K2_STEPS : part * 8 + 8;
#pragma unroll
                    for (int k = k_lo; k < k_hi; ++k)
                        tc::mma_f16_ss_pred(tmem + TM_W1, tc::smem_desc(XS3_MN, xs3 + k * 256), tc::smem_desc(DZ1_MN, p1 + k * 16 * 128), ID_WG1, (it | k) != 0, leader);
                };
                // conv2 wgrad K steps + dgrad of P1-row tile `tile`
                auto bwd2_tile = [&](int it, int tile) {
                    // conv2 wgrad, both vertical taps in ONE MMA per K step: dW2[ty][tx][k][oc] = sum_R' P1[R' + tx][k] * DZ2[R' - 13 ty][oc], so with the A view fixed
                    // the two ty's are two N atoms of B thirteen rows apart (columns [0,32) = ty 1, [32,64) = ty 0; the zero halo in front of DZ2 and the zero
                    // rows behind P1 make the out-of-range products vanish).  K runs over R' = 0 .. 447; tile t may issue the steps whose DZ2 rows exist.
                    const int k_lo = tile * 8, k_hi = tile == 3 ? K2_STEPS + 1 : tile * 8 + 8;
#pragma unroll
                    for (int k = k_lo; k < k_hi; ++k)
                        tc::mma_f16_ss_pred(tmem + TM_W2, tc::smem_desc(P1_MN, p1 + k * 16 * 128), tc::smem_desc(DZ2_MN2, dz2 + (k * 16 - P1_W) * 64), ID_WG2, (it | k) != 0,
                                            leader);
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            tc::mma_f16_ss_pred(tmem + TM_DGR + (tile & 1) * 64, tc::smem_desc(KM64, dz2 + (tile * 128 - ((t >> 1) * P1_W + (t & 1))) * 64 + k * 32),
                                                tc::smem_desc(KM64, w2d + t * 4096 + k * 32), ID_DG, (t | k) != 0, leader);
                    tc::mma_commit_pred(&dg_done[tile], leader);
                };
                // prologue: conv1 of the first sample
                tc::mbar_wait(&xs_ready, 0);
                tc::tc_fence_after();
#pragma unroll
                for (int tile = 0; tile < 8; ++tile) conv1_tile(tile);
                tc::mma_commit_pred(&c1a_done, leader);
#pragma unroll
                for (int tile = 8; tile < 14; ++tile) conv1_tile(tile);
                tc::mma_commit_pred(&c1b_done, leader);
                for (int it = 0; it < n_my; ++it) {
                    const uint32_t ph = (uint32_t)(it & 1);
                    PASS_TS(16);
                    // ---- C2, tile by tile behind the conv1 epilogue
#pragma unroll
                    for (int tile = 0; tile < 4; ++tile) {
                        tc::mbar_wait(&p1_ready[tile], ph);
                        tc::tc_fence_after();
                        if (BIAS_MMA) tc::mma_f16_ss_pred(tmem + TM_X + tile * 32, tc::smem_desc(ONES_K, onesa), tc::smem_desc(BB_K, bb2), ID_C2, 0, leader);
#pragma unroll
                        for (int t = 0; t < 4; ++t)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc::mma_f16_ss_pred(tmem + TM_X + tile * 32, tc::smem_desc(KM128, p1 + (tile * 128 + (t >> 1) * P1_W + (t & 1)) * 128 + k * 32),
                                                    tc::smem_desc(KM128, w2 + t * 4096 + k * 32), ID_C2, (BIAS_MMA || (t | k) != 0) ? 1u : 0u, leader);
                        tc::mma_commit_pred(&c2_done[tile], leader);
                    }
                    PASS_DBG(1, it * 16 + 2);
                    PASS_TS(17);
                    // ---- conv1 of the NEXT sample, tiles 0..7, in the window where the tensor pipe would wait for the loss and the DZ2 pass: its XS rows were
                    // built under this sample's conv2, its accumulator columns held this sample's conv2 accumulators (read by the conv2 epilogue: c2_read)
                    if (it + 1 < n_my) {
                        tc::mbar_wait(&xs_ready, ph ^ 1u);
                        tc::mbar_wait(&c2_read, ph);
                        tc::tc_fence_after();
#pragma unroll
                        for (int tile = 0; tile < 8; ++tile) conv1_tile(tile);
                        tc::mma_commit_pred(&c1a_done, leader);
                    }
                    // ---- backward, conv2: wgrad K steps + dgrad tile, tile by tile behind the DZ2 pass (dgrad ring: tile t + 2 after the epilogue has read tile t)
#pragma unroll
                    for (int tile = 0; tile < 4; ++tile) {
                        tc::mbar_wait(&dz2_ready[tile], ph);
                        if (tile >= 2) tc::mbar_wait(&dg_read[tile - 2], ph);
                        tc::tc_fence_after();
                        bwd2_tile(it, tile);
                    }
                    PASS_DBG(1, it * 16 + 3);
                    PASS_TS(18);
                    // ---- backward, conv1 wgrad: the K steps of a P1-row tile as soon as its DZ1 rows exist.  Parts 0..2 (DZ1 rows < 384) get their own commit:
                    // the next sample's conv1 epilogue may overwrite those rows while part 3 and conv1 tiles 8..13 are still in the pipe
                    tc::mbar_wait(&full_xs3, ph);
#pragma unroll
                    for (int part = 0; part < 3; ++part) {
                        tc::mbar_wait(&dz1_ready[part], ph);
                        tc::tc_fence_after();
                        wgrad1_part(it, part);
                    }
                    tc::mma_commit_pred(&wg1a_done, leader);
                    PASS_TS(19);
                    tc::mbar_wait(&dz1_ready[3], ph);
                    tc::tc_fence_after();
                    wgrad1_part(it, 3);
                    tc::mma_commit_pred(&mma2_done, leader);
                    PASS_DBG(1, it * 16 + 4);
                    PASS_TS(20);
                    // ---- conv1 of the next sample, tiles 8..13: the dgrad ring they overwrite has been read (dz1_ready[3])
                    if (it + 1 < n_my) {
#pragma unroll
                        for (int tile = 8; tile < 14; ++tile) conv1_tile(tile);
                        tc::mma_commit_pred(&c1b_done, leader);
                    }
                    PASS_TS(21);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ workers: thread (q, lane, g) owns TMEM lane q*32+lane, column group g
        const int q = warp & 3, g = warp >> 2, tl = q * 32 + lane, w = threadIdx.x;
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        const uint32_t xs_s = tc::smem_u32(smem + SM_XS), p1_s = tc::smem_u32(smem + SM_P1), x_s = tc::smem_u32(smem + SM_X), dz2s = tc::smem_u32(smem + SM_DZ2) + 1024;
        const uint32_t b1_s = tc::smem_u32(bias_s) + (g & 1) * 32, b2_s = tc::smem_u32(bias_s + 16) + g * 32, lp_s = tc::smem_u32(&logit_part[0][0]);
        float db2[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) db2[c] = 0.f;
        float dbfc = 0.f;
        double loss_acc = 0.0;
        const float bfc = a.bfc[0];
        {   // fc.weight gradient accumulators start at zero
            uint32_t zr[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) zr[c] = 0u;
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) tmem_st_32x8(tmem + tlane + TM_FC + tile * 32 + g * 8, zr);
            tmem_st_wait();
        }
        // fc.weight slice of this thread's rows and columns (the same for every sample): kept in registers; used by the forward dot and by DZ2
        float wreg[4][8];
#pragma unroll
        for (int tile = 0; tile < 4; ++tile) {
            const int R = tile * 128 + tl;
#pragma unroll
            for (int c4 = 0; c4 < 2; ++c4) {
                const float4 wv = R < P1_ROWS ? reinterpret_cast<const float4*>(a.wfcp + R * 32 + g * 8)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
                wreg[tile][4 * c4] = wv.x; wreg[tile][4 * c4 + 1] = wv.y; wreg[tile][4 * c4 + 2] = wv.z; wreg[tile][4 * c4 + 3] = wv.w;
            }
        }
        // ---- XSb: XS rows (8 values = (dy,dx,ch) of super pixel (sy,sx) of the zero-padded input)
        auto build_xs = [&](int it) {
            if (!a.x_f32) {
                tc::mbar_wait(&x_full, (uint32_t)(it & 1));
                // all of this thread's byte loads (4 rows x 8) are issued before any is used: one shared-memory latency instead of four
                constexpr int NR = (XS_ROWS + NWORK - 1) / NWORK;
                uint32_t u[NR][8];
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int rr = w + i * NWORK;
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
                    const int rr = w + i * NWORK;
                    if (rr >= XS_ROWS) break;
                    uint32_t f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __float_as_uint((float)u[i][e]);      // integers 0..255 are exact in bf16 = the high half of the float
                    tc::sts128(xs_s + rr * 16, make_uint4(__byte_perm(f[0], f[1], 0x7632), __byte_perm(f[2], f[3], 0x7632), __byte_perm(f[4], f[5], 0x7632),
                                                          __byte_perm(f[6], f[7], 0x7632)));
                }
            } else {
                const int b = blockIdx.x + it * gridDim.x;
                long long xb = a.x_index ? a.x_index[b] : b;
                if (a.x_rows > 0 && (unsigned long long)xb >= (unsigned long long)a.x_rows) { *a.oob = 1; xb = 0; }
                for (int rr = w; rr < XS_ROWS; rr += NWORK) {
                    const int sy = rr / XS_W, sx = rr - sy * XS_W;
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int iy = 2 * sy + (e >> 2) - 1, ix = 2 * sx + ((e >> 1) & 1) - 1, ch = e & 1;
                        v[e] = (iy >= 0 && iy < 128 && ix >= 0 && ix < 50) ? reinterpret_cast<const float*>(a.x)[(size_t)xb * 12800 + (ch * 128 + iy) * 50 + ix] : 0.f;
                    }
                    tc::sts128(xs_s + rr * 16, make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
                }
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&xs_ready);
        };
        if (n_my > 0) build_xs(0);
        for (int it = 0; it < n_my; ++it) {
            const int b = blockIdx.x + it * gridDim.x;
            const uint32_t ph = (uint32_t)(it & 1);
            // ---- S3: conv1 epilogue -> P1 (rows = super pixels, 64 values = (dy,dx,c16)); tile parity g >> 1, 8 of the 16 channels (g & 1) per thread.
            // conv1 tiles 0..7 of this sample were issued in the middle of the previous sample's backward; the P1 rows are free once that sample's
            // conv1 wgrad MMAs have read DZ1.
            if (w == 0) PASS_TS(0);
            if (it > 0) tc::mbar_wait(&wg1a_done, ph ^ 1u);           // DZ1 rows < 384 of the previous sample have been read (conv1 tiles <= 10 write P1 rows <= 363)
            tc::mbar_wait(&c1a_done, ph);
            tc::tc_fence_after();
            if (w == 0) { PASS_DBG(2, it * 16 + 1); PASS_TS(1); }
#pragma unroll
            for (int tt = 0; tt < 7; ++tt) {
                const int tile = 2 * tt + (g >> 1);
                if (tt == 4) {                                        // tiles 8..13 were issued behind the previous sample's last conv1 wgrad part (DZ1 rows >= 384)
                    if (it > 0) tc::mbar_wait(&mma2_done, ph ^ 1u);
                    tc::mbar_wait(&c1b_done, ph);
                    tc::tc_fence_after();
                }
                uint32_t r[8];
                tmem_ld_32x8(tmem + tlane + TM_X + tile * 16 + (g & 1) * 8, r);
                tc::tmem_ld_wait();
                const int m = tile * 128 + tl, oy = m / XS_W, ox = m - oy * XS_W;
                if (m < XS_ROWS && oy < 64 && ox < 25) {
                    float z[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) z[c] = __uint_as_float(r[c]);
                    if (!BIAS_MMA) {
                        const uint4 bq0 = tc::lds128(b1_s), bq1 = tc::lds128(b1_s + 16);
                        z[0] += __uint_as_float(bq0.x); z[1] += __uint_as_float(bq0.y); z[2] += __uint_as_float(bq0.z); z[3] += __uint_as_float(bq0.w);
                        z[4] += __uint_as_float(bq1.x); z[5] += __uint_as_float(bq1.y); z[6] += __uint_as_float(bq1.z); z[7] += __uint_as_float(bq1.w);
                    }
                    uint32_t o[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[c] = pack_bf16x2(fmaxf(z[2 * c], 0.2f * z[2 * c]), fmaxf(z[2 * c + 1], 0.2f * z[2 * c + 1]));      // LeakyReLU(0.2) = max(z, 0.2 z)
                    const int yp = oy + 1, xp = ox + 1, R = (yp >> 1) * P1_W + (xp >> 1), cell = (yp & 1) * 2 + (xp & 1);
                    tc::sts128(p1_s + R * 128 + (((2 * cell + (g & 1)) ^ (R & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
                }
                // conv1 tiles <= 2 tt + 1 are in P1 once every worker is past this point: conv2 tile 0 reads P1 rows <= 141
                // (conv1 rows <= 544: tiles 0..4), tile 1 rows <= 269 (tiles 0..8), tiles 2 and 3 everything
                if (tt == 2 || tt == 4 || tt == 6) {
                    tc::tc_fence_before();
                    tc::fence_proxy_async_smem();
                    tc::mbar_arrive(&p1_ready[tt == 2 ? 0 : tt == 4 ? 1 : 2]);
                    if (tt == 6) tc::mbar_arrive(&p1_ready[3]);
                }
            }
            if (w == 0) PASS_TS(2);
            // ---- XSb of the next sample, under this sample's conv2 MMAs (every conv1 MMA of this sample has read the XS rows: c1b_done)
            if (it + 1 < n_my) {
                tc::mbar_wait(&xs_saved, ph);                         // the producer's copy of XS(it) has read the rows
                build_xs(it + 1);
            }
            if (w == 0) PASS_TS(3);
            // ---- S5: conv2 epilogue -> A2 rows (64-byte swizzle, where DZ2 will be), fc partial dot; 8 of the 32 channels (g) per thread
            float dot = 0.f;
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {
                tc::mbar_wait(&c2_done[tile], ph);
                tc::tc_fence_after();
                const int R = tile * 128 + tl;
                uint32_t r[8];
                tmem_ld_32x8(tmem + tlane + TM_X + tile * 32 + g * 8, r);
                tc::tmem_ld_wait();
                if (R >= P1_ROWS) continue;
                const int oy = R / P1_W, ox = R - oy * P1_W;
                const bool real = oy < 32 && ox < 12;                 // junk rows of the row space are stored as zeros
                float z[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) z[c] = __uint_as_float(r[c]);
                if (!BIAS_MMA) {
                    const uint4 bq0 = tc::lds128(b2_s), bq1 = tc::lds128(b2_s + 16);
                    z[0] += __uint_as_float(bq0.x); z[1] += __uint_as_float(bq0.y); z[2] += __uint_as_float(bq0.z); z[3] += __uint_as_float(bq0.w);
                    z[4] += __uint_as_float(bq1.x); z[5] += __uint_as_float(bq1.y); z[6] += __uint_as_float(bq1.z); z[7] += __uint_as_float(bq1.w);
                }
                uint32_t o[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    o[c] = real ? pack_bf16x2(fmaxf(z[2 * c], 0.2f * z[2 * c]), fmaxf(z[2 * c + 1], 0.2f * z[2 * c + 1])) : 0u;
                    dot = fmaf(bf_lo(o[c]), wreg[tile][2 * c], dot);                      // the bf16 values the backward reads
                    dot = fmaf(bf_hi(o[c]), wreg[tile][2 * c + 1], dot);
                }
                tc::sts128(dz2s + R * 64 + ((g ^ ((R >> 1) & 3)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
            }
            dot = warp_sum(dot);
            if (lane == 0) logit_part[ph][warp] = dot;
            tc::tc_fence_before();
            tc::mbar_arrive(&c2_read);                                // the conv2 accumulator columns may take conv1 tiles 0..7 of the next sample
            if (w == 0) PASS_TS(4);
            bar_workers();
            // ---- BCE with logits (network_tests.py:304-306,313): loss_b = max(x,0) - x y + log1p(exp(-|x|)); dlogit = (sigmoid(x) - y) / n.
            // The sixteen partial dots are added in a fixed order: the logit does not depend on which warp finished first.
            float xl = bfc;
            {
                const uint32_t lp = lp_s + ph * 64;
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const uint4 pq = tc::lds128(lp + i4 * 16);
                    xl += __uint_as_float(pq.x); xl += __uint_as_float(pq.y); xl += __uint_as_float(pq.z); xl += __uint_as_float(pq.w);
                }
            }
            const float dl = (1.f / (1.f + expf(-xl)) - a.y) * a.inv_n;
            if (w == 0) {
                if (a.logits) a.logits[b] = xl;
                loss_acc += (double)(fmaxf(xl, 0.f) - xl * a.y + log1pf(expf(-fabsf(xl))));
                dbfc += dl;
                PASS_DBG(2, it * 16 + 2);
            }
            // ---- W1: A2 -> DZ2 in place, fc.weight / conv2.bias gradients
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {
                const int R = tile * 128 + tl;
                if (tile * 128 + q * 32 >= A2_ROWS) {                 // warp-uniform: this warp's 32 rows are all beyond the tile's rows
                    tc::mbar_arrive(&dz2_ready[tile]);
                    continue;
                }
                const bool inb = R < A2_ROWS;                         // rows 429..431 have w = 0 -> they are (re)written as zeros
                const uint32_t cp = dz2s + (inb ? R : 0) * 64 + ((g ^ ((R >> 1) & 3)) << 4);
                const uint4 av = tc::lds128(cp);
                const uint32_t au[4] = {av.x, av.y, av.z, av.w};
                uint32_t acc[8], o[4];
                tmem_ld_32x8(tmem + tlane + TM_FC + tile * 32 + g * 8, acc);          // .sync.aligned: every lane of the warp takes part
                tc::tmem_ld_wait();
                const bool real = R < P1_ROWS;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float x0 = real ? bf_lo(au[j]) : 0.f, x1 = real ? bf_hi(au[j]) : 0.f;
                    float g0 = dl * wreg[tile][2 * j], g1 = dl * wreg[tile][2 * j + 1];
                    if (!(x0 > 0.f)) g0 *= 0.2f;
                    if (!(x1 > 0.f)) g1 *= 0.2f;
                    o[j] = pack_bf16x2(g0, g1);
                    db2[2 * j] += g0; db2[2 * j + 1] += g1;
                    acc[2 * j] = __float_as_uint(fmaf(dl, x0, __uint_as_float(acc[2 * j])));
                    acc[2 * j + 1] = __float_as_uint(fmaf(dl, x1, __uint_as_float(acc[2 * j + 1])));
                }
                tmem_st_32x8(tmem + tlane + TM_FC + tile * 32 + g * 8, acc);
                if (inb) tc::sts128(cp, make_uint4(o[0], o[1], o[2], o[3]));
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&dz2_ready[tile]);                    // the MMAs of this tile may start
            }
            tmem_st_wait();
            if (w == 0) PASS_TS(5);
            // ---- W3: conv2 dgrad epilogue -> DZ1 in place over P1 (tile t as soon as its MMAs have committed); cell g = (dy, dx) per thread.
            // The conv1.bias gradient (column sums of DZ1) comes out of the conv1 wgrad MMAs (ones plane); those run tile by tile behind this loop.
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {
                tc::mbar_wait(&dg_done[tile], ph);
                tc::tc_fence_after();
                const int R = tile * 128 + tl;
                uint32_t r[16];
                tc::tmem_ld_32x16(tmem + tlane + TM_DGR + (tile & 1) * 64 + g * 16, r);
                tc::tmem_ld_wait();
                if (tile < 2) {                                       // ring slot `tile` may take dgrad tile + 2
                    tc::tc_fence_before();
                    tc::mbar_arrive(&dg_read[tile]);
                }
                if (R < A2_ROWS) {
                    const int sy = R / P1_W, sx = R - sy * P1_W;
                    const int oy = 2 * sy + (g >> 1) - 1, ox = 2 * sx + (g & 1) - 1;
                    // zero-padding cells of P1 have no conv1 output behind them; rows 429..431 are pad rows: both become 0
                    const bool cell = R < P1_ROWS && oy >= 0 && oy < 64 && ox >= 0 && ox < 25;
                    const uint32_t prow = p1_s + R * 128;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {                  // 8 channels per 16-byte chunk
                        const uint32_t addr = prow + (((g * 2 + hh) ^ (R & 7)) << 4);
                        uint32_t o[4] = {0u, 0u, 0u, 0u};
                        if (cell) {
                            const uint4 av = tc::lds128(addr);
                            const uint32_t au[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float g0 = __uint_as_float(r[hh * 8 + 2 * j]), g1 = __uint_as_float(r[hh * 8 + 2 * j + 1]);
                                if (!pos_lo(au[j])) g0 *= 0.2f;
                                if (!pos_hi(au[j])) g1 *= 0.2f;
                                o[j] = pack_bf16x2(g0, g1);
                            }
                        }
                        tc::sts128(addr, make_uint4(o[0], o[1], o[2], o[3]));
                    }
                }
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&dz1_ready[tile]);                    // DZ1 rows of this tile are complete: their conv1 wgrad K steps may be issued
            }
            if (w == 0) { PASS_DBG(2, it * 16 + 3); PASS_TS(6); }
        }
        // ------------------------------------------------------------------ flush: weight gradients leave the SM once per launch
        if (n_my > 0) {
            tc::mbar_wait(&mma2_done, (uint32_t)((n_my - 1) & 1));
            tc::tc_fence_after();
            {   // conv2: TMEM lane m = tx*64 + (dy*2+dx)*16 + ic, column = (1-ty)*32 + oc  ->  conv2.weight[oc][ic][2ty+dy][2tx+dx]; this thread: 16 oc of one ty
                const int tx = tl >> 6, dy = (tl >> 5) & 1, dx = (tl >> 4) & 1, ic = tl & 15, ty = 1 - (g >> 1), oc0 = (g & 1) * 16;      // columns [0,32) hold ty = 1
                uint32_t r[16];
                tc::tmem_ld_32x16(tmem + tlane + TM_W2 + g * 16, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) atomicAdd(&a.dw2[(((oc0 + c) * 16 + ic) * 4 + 2 * ty + dy) * 4 + 2 * tx + dx], __uint_as_float(r[c]));
            }
            {   // conv1: TMEM lane m = patch value ((ay*3+ax)*8 + (dy',dx',ch)), column = cell*16 + oc; patch pixel (py,px) = (2ay+dy', 2ax+dx')
                // feeds cell (dy,dx) through tap (ky,kx) = (py - 2dy, px - 2dx)  ->  conv1.weight[oc][ch][ky][kx]; lanes 72..79: sum_R DZ1[R][cell*16+oc]
                uint32_t r[16];
                tc::tmem_ld_32x16(tmem + tlane + TM_W1 + g * 16, r);                  // cell g
                tc::tmem_ld_wait();
                if (tl < 72) {
                    const int at = tl >> 3, e = tl & 7, py = 2 * (at / 3) + (e >> 2), px = 2 * (at % 3) + ((e >> 1) & 1), ch = e & 1;
                    const int ky = py - 2 * (g >> 1), kx = px - 2 * (g & 1);
                    if (ky >= 0 && ky <= 3 && kx >= 0 && kx <= 3) {
#pragma unroll
                        for (int oc = 0; oc < 16; ++oc) atomicAdd(&a.dw1[((oc * 2 + ch) * 4 + ky) * 4 + kx], __uint_as_float(r[oc]));
                    }
                } else if (tl == 72) {
#pragma unroll
                    for (int oc = 0; oc < 16; ++oc) atomicAdd(&red_s[32 + oc], __uint_as_float(r[oc]));
                }
            }
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {                    // fc.weight[0, oc*384 + oy*12 + ox]
                const int R = tile * 128 + tl;
                uint32_t r[8];
                tmem_ld_32x8(tmem + tlane + TM_FC + tile * 32 + g * 8, r);
                tc::tmem_ld_wait();
                const int oy = R / P1_W, ox = R - oy * P1_W;
                if (R < P1_ROWS && oy < 32 && ox < 12) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) atomicAdd(&a.dwfc[(g * 8 + c) * 384 + oy * 12 + ox], __uint_as_float(r[c]));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float s2 = warp_sum(db2[c]);
            if (lane == 0) atomicAdd(&red_s[g * 8 + c], s2);
        }
        if (w == 0 && n_my > 0) {
            atomicAdd(a.dbfc, dbfc);
            if (a.loss) atomicAdd(a.loss, (float)(loss_acc * (double)a.inv_n));
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc(tmem, 512);
    if (threadIdx.x < 32) atomicAdd(&a.db2[threadIdx.x], red_s[threadIdx.x]);
    else if (threadIdx.x < 48) atomicAdd(&a.db1[threadIdx.x - 32], red_s[threadIdx.x]);
}

}  // namespace

static int g_pass_flags = 0;

extern "C" {

// bit 0: biases added in the epilogues instead of through the tensor pipe (one extra MMA per tile, A = ones, B = bias hi | lo).  Process-wide; for A/B checks.
int mmg_disc_pass_set_flags(int flags) { g_pass_flags = flags; return MMG_OK; }

size_t mmg_disc_pass_workspace_bytes(void) { return (size_t)MMG_NUM_SMS * 2 * XS_ROWS * 16 + 128; }      // scratch slots + the out-of-range flag

// x (B,2,128,50) uint8 (x_dtype 2) or float32 (0), optionally gathered through x_index (int64, device; checked against x_rows when > 0: an
// out-of-range index reads row 0 and sets the int32 flag in the last 128 bytes of the workspace).  BCE target `target`, loss mean and
// dlogit over `loss_rows` rows (0 = B).  logits (B,) and loss[0] (+=) are optional; the six fp32 gradients are ACCUMULATED (+=).
int mmg_disc_pass_fused_dbg(const void* x, int x_dtype, const int64_t* x_index, int64_t x_rows, const void* packed, const float* conv1_b, const float* conv2_b, const float* fc_b,
                            float target, int64_t loss_rows, float* logits, float* loss, float* dconv1_w, float* dconv1_b, float* dconv2_w, float* dconv2_b,
                            float* dfc_w, float* dfc_b, void* workspace, size_t ws_bytes, int64_t B, void* stream, int* dbg) {
    MMG_REQUIRE(x && packed && conv1_b && conv2_b && fc_b && dconv1_w && dconv1_b && dconv2_w && dconv2_b && dfc_w && dfc_b && workspace && B >= 0 && loss_rows >= 0,
                MMG_EINVAL, "disc_pass_fused: bad arguments");
    MMG_REQUIRE(x_dtype == 0 || x_dtype == 2, MMG_EINVAL, "disc_pass_fused: x_dtype must be 0 (f32) or 2 (u8)");
    MMG_REQUIRE(ws_bytes >= mmg_disc_pass_workspace_bytes(), MMG_EWORKSPACE, "disc_pass_fused: workspace too small");
    if (B == 0) return MMG_OK;
    MMG_REQUIRE(B < (1LL << 31) / XS_ROWS, MMG_EUNSUPPORTED, "disc_pass_fused: batch too large");
    MMG_REQUIRE(x_dtype != 2 || ((uintptr_t)x & 15) == 0, MMG_EINVAL, "disc_pass_fused: x must be 16-byte aligned");
    MMG_REQUIRE(((uintptr_t)workspace & 127) == 0, MMG_EINVAL, "disc_pass_fused: workspace must be 128-byte aligned");
    const unsigned char* pk = (const unsigned char*)packed;
    const int grid = (int)(B < MMG_NUM_SMS ? B : MMG_NUM_SMS);
    CUtensorMap map_xs3, map_w1, map_w2, map_w2d;
    {   // the scratch slots as (slots, 65, 26, 8): every second super pixel in both directions -> one 33 x 13 plane of 16-byte rows per load
        const uint64_t dims[4] = {8, (uint64_t)XS_W, 65, (uint64_t)(2 * grid)}, strides[3] = {16, 16 * XS_W, 16 * (uint64_t)XS_ROWS};
        const uint32_t box[4] = {8, 26, 65, 1}, estr[4] = {1, 2, 2, 1};
        MMG_REQUIRE(tc::make_map_nd_bf16(&map_xs3, workspace, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE) == 0, MMG_EINVAL, "disc_pass_fused: tensor map (xs3)");
    }
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w1, pk, 16, 32, 32, 16, 32, CU_TENSOR_MAP_SWIZZLE_32B) == 0, MMG_EINVAL, "disc_pass_fused: tensor map (w1b)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w2, pk + 2048, 64, 128, 128, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B) == 0, MMG_EINVAL, "disc_pass_fused: tensor map (w2p)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w2d, pk + 2048 + 16384, 32, 256, 64, 32, 256, CU_TENSOR_MAP_SWIZZLE_64B) == 0, MMG_EINVAL, "disc_pass_fused: tensor map (w2d)");
    PassArgs a;
    a.x = x; a.x_f32 = x_dtype == 0; a.x_index = x_index; a.x_rows = x_index ? x_rows : 0; a.oob = (int*)((unsigned char*)workspace + (size_t)MMG_NUM_SMS * 2 * XS_ROWS * 16);
    a.b1 = conv1_b; a.b2 = conv2_b; a.wfcp = (const float*)(pk + 2048 + 32768); a.bfc = fc_b;
    a.y = target; a.inv_n = 1.f / (float)(loss_rows ? loss_rows : B); a.logits = logits; a.loss = loss;
    a.dw1 = dconv1_w; a.db1 = dconv1_b; a.dw2 = dconv2_w; a.db2 = dconv2_b; a.dwfc = dfc_w; a.dbfc = dfc_b;
    a.scratch = (__nv_bfloat16*)workspace; a.B = (int)B; a.dbg = dbg;
    if (g_pass_flags & 1) {         // bit 0 of mmg_disc_pass_set_flags: biases added in the epilogues instead of by one extra MMA per tile (measured 5 % slower: the conv1
                                    // epilogue is on the critical path, the bias MMAs are not)
        MMG_CUDA(cudaFuncSetAttribute(disc_pass_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL + 1024));
        disc_pass_fused_kernel<false><<<grid, NTHREADS, SM_TOTAL + 1024, (cudaStream_t)stream>>>(map_xs3, map_w1, map_w2, map_w2d, a);
    } else {
        MMG_CUDA(cudaFuncSetAttribute(disc_pass_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL + 1024));
        disc_pass_fused_kernel<true><<<grid, NTHREADS, SM_TOTAL + 1024, (cudaStream_t)stream>>>(map_xs3, map_w1, map_w2, map_w2d, a);
    }
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

int mmg_disc_pass_fused(const void* x, int x_dtype, const int64_t* x_index, int64_t x_rows, const void* packed, const float* conv1_b, const float* conv2_b, const float* fc_b,
                        float target, int64_t loss_rows, float* logits, float* loss, float* dconv1_w, float* dconv1_b, float* dconv2_w, float* dconv2_b,
                        float* dfc_w, float* dfc_b, void* workspace, size_t ws_bytes, int64_t B, void* stream) {
    return mmg_disc_pass_fused_dbg(x, x_dtype, x_index, x_rows, packed, conv1_b, conv2_b, fc_b, target, loss_rows, logits, loss, dconv1_w, dconv1_b, dconv2_w, dconv2_b, dfc_w,
                                   dfc_b, workspace, ws_bytes, B, stream, nullptr);
}

}  // extern "C"
