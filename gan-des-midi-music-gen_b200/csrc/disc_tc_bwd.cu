// bf16 tensor-core path of the MM-GAN discriminator, backward (autograd of network_tests.py:156-160,
// reached by disc_loss.backward() / gen_loss.backward(), network_tests.py:307,314).  Layouts: see disc_tc.cu.
//
//   fc_bwd      (SIMT, HBM-bound)  dz2 = dlogit[b] * fc.w * lrelu'(a2)  -> DZ2 bf16;  dfc.w += dlogit[b]*a2;  dconv2.b += dz2
//   conv2 wgrad (tcgen05)          dW2[t][k64][oc] = sum_rows P1[row+shift_t][k] * DZ2[row][oc]    both operands MN-major:
//                                  M = 128 = the two horizontally adjacent taps (second atom = same box, one row later)
//   conv2 dgrad (tcgen05)          da1[R][n64] = sum_t DZ2[R-shift_t][oc] * W2d[t][n][oc];  epilogue: dz1 = da1*lrelu'(a1),
//                                  dconv1.b += dz1, DZ1c bf16 written in conv1's own row space (B*1690, 16)
//   conv1 wgrad: csrc/disc_tc_conv1.cu (tcgen05)
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int ROWS_PER_SAMPLE = 429;
constexpr int SG_W = 13;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------
// fc backward + LeakyReLU' : thread = (row-in-sample rr, 8 channels), loops over a slice of samples
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fc_bwd_kernel(const __nv_bfloat16* __restrict__ a2, const float* __restrict__ dlogit,
                                                     const float* __restrict__ wfcp, __nv_bfloat16* __restrict__ dz2, float* __restrict__ dwfc,
                                                     float* __restrict__ db2, int B, int b_per_block) {
    const int rr = blockIdx.x * 64 + (threadIdx.x >> 2), ch = (threadIdx.x & 3) * 8;
    const int b0 = blockIdx.y * b_per_block, b1 = min(B, b0 + b_per_block);
    __shared__ float db_s[32];
    if (threadIdx.x < 32) db_s[threadIdx.x] = 0.f;
    __syncthreads();
    float dw[8], dbv[8], w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dw[j] = dbv[j] = w[j] = 0.f;
    const bool in = rr < ROWS_PER_SAMPLE;
    if (in) {
        const float4* wp = reinterpret_cast<const float4*>(wfcp + rr * 32 + ch);
        const float4 w0 = wp[0], w1 = wp[1];
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
        for (int b = b0; b < b1; ++b) {
            const size_t off = ((size_t)b * ROWS_PER_SAMPLE + rr) * 32 + ch;
            const uint4 av = *reinterpret_cast<const uint4*>(a2 + off);
            const float dl = dlogit[b];
            const uint32_t au[4] = {av.x, av.y, av.z, av.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x0 = bf_lo(au[j]), x1 = bf_hi(au[j]);
                const float g0 = dl * w[2 * j] * (x0 > 0.f ? 1.f : 0.2f), g1 = dl * w[2 * j + 1] * (x1 > 0.f ? 1.f : 0.2f);
                o[j] = pack_bf16x2(g0, g1);
                dw[2 * j] = fmaf(dl, x0, dw[2 * j]); dw[2 * j + 1] = fmaf(dl, x1, dw[2 * j + 1]);
                dbv[2 * j] += bf_lo(o[j]); dbv[2 * j + 1] += bf_hi(o[j]);          // what conv2 wgrad/dgrad will read
            }
            *reinterpret_cast<uint4*>(dz2 + off) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        const int oy = rr / SG_W, ox = rr - oy * SG_W;
        const bool real = oy < 32 && ox < 12;                      // junk rows carry no fc weight
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (real) atomicAdd(&dwfc[(ch + j) * 384 + oy * 12 + ox], dw[j]);     // fc.weight[0, oc*384 + oy*12 + ox]
            atomicAdd(&db_s[ch + j], dbv[j]);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) atomicAdd(&db2[threadIdx.x], db_s[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// conv2 weight gradient (tcgen05, both operands MN-major, accumulators live in TMEM for the CTA's whole slice)
// ------------------------------------------------------------------------------------------------
constexpr int WG_STAGES = 4;
constexpr int WG_A_BYTES = 144 * 128;                  // P1 box: 128 rows + 14 halo (+2)
constexpr int WG_B_BYTES = 128 * 64;                   // DZ2 box
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024;

__global__ void __launch_bounds__(192, 1) conv2_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_p1, const __grid_constant__ CUtensorMap map_dz2,
                                                                float* __restrict__ dw2, int num_chunks) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full[WG_STAGES], empty[WG_STAGES], done;
    __shared__ uint32_t tmem_s;
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c_lo = (int)((long long)num_chunks * blockIdx.x / gridDim.x), c_hi = (int)((long long)num_chunks * (blockIdx.x + 1) / gridDim.x);
    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_s, 64); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;
    if (c_hi > c_lo) {
        if (warp == 0 && tc::elect_one()) {
            for (int c = c_lo, it = 0; c < c_hi; ++c, ++it) {
                const int stage = it % WG_STAGES, phase = (it / WG_STAGES) & 1;
                tc::mbar_wait(&empty[stage], phase ^ 1);
                tc::mbar_expect_tx(&full[stage], WG_STAGE_BYTES);
                unsigned char* st = smem + stage * WG_STAGE_BYTES;
                tc::tma_load_2d(st, &map_p1, &full[stage], 0, c * 128);
                tc::tma_load_2d(st + WG_A_BYTES, &map_dz2, &full[stage], 0, c * 128);
            }
        } else if (warp == 1 && tc::elect_one()) {
            constexpr uint64_t A_MN = tc::smem_desc_base(128, 1024, tc::SW_128B);     // atom 1 = one row (128 B) later, 8-row K groups 1024 B apart
            constexpr uint64_t B_MN = tc::smem_desc_base(0, 512, tc::SW_64B);
            constexpr uint32_t IDESC = tc::idesc_bf16(128, 32, 1, 1);
            for (int c = c_lo, it = 0; c < c_hi; ++c, ++it) {
                const int stage = it % WG_STAGES, phase = (it / WG_STAGES) & 1;
                tc::mbar_wait(&full[stage], phase);
                tc::tc_fence_after();
                const uint32_t a_base = tc::smem_u32(smem + stage * WG_STAGE_BYTES), b_base = a_base + WG_A_BYTES;
#pragma unroll
                for (int ty = 0; ty < 2; ++ty)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        tc::mma_f16_ss(tmem + ty * 32, tc::smem_desc(A_MN, a_base + (ty * SG_W) * 128 + k * 16 * 128),
                                       tc::smem_desc(B_MN, b_base + k * 16 * 64), IDESC, (it | k) != 0);
                tc::mma_commit(&empty[stage]);
            }
            tc::mma_commit(&done);
        } else if (warp >= 2) {
            const int q = warp & 3;
            tc::mbar_wait(&done, 0);
            tc::tc_fence_after();
            // TMEM lane m = tx*64 + (dy*2+dx)*16 + ic, column = ty*32 + oc  ->  conv2.weight[oc][ic][2ty+dy][2tx+dx]
            const int m = q * 32 + lane, tx = m >> 6, dy = (m >> 5) & 1, dx = (m >> 4) & 1, ic = m & 15;
#pragma unroll
            for (int ty = 0; ty < 2; ++ty) {
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + ty * 32, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int oc = 0; oc < 32; ++oc)
                    atomicAdd(&dw2[((oc * 16 + ic) * 4 + 2 * ty + dy) * 4 + 2 * tx + dx], __uint_as_float(r[oc]));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------
// conv2 data gradient (tcgen05 tap-shift over DZ2) fused with LeakyReLU' of conv1 and the conv1 bias gradient
// ------------------------------------------------------------------------------------------------
constexpr int DG_STAGES = 4;
constexpr int DG_A_BYTES = 144 * 64;
constexpr int DG_W_BYTES = 4 * 64 * 64;
constexpr int DG_SMEM = DG_W_BYTES + DG_STAGES * DG_A_BYTES + 1024;

__global__ void __launch_bounds__(192, 2) conv2_dgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dz2, const __grid_constant__ CUtensorMap map_w,
                                                                const __nv_bfloat16* __restrict__ p1, __nv_bfloat16* __restrict__ dz1,
                                                                float* __restrict__ db1, int total_rows, int num_tiles) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full[DG_STAGES], empty[DG_STAGES], tfull[2], tempty[2], wbar;
    __shared__ uint32_t tmem_s;
    __shared__ float db_s[16];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* smem_w = smem;
    unsigned char* smem_a = smem + DG_W_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < DG_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
        tc::mbar_init(&wbar, 1);
        tc::fence_barrier_init();
    }
    if (threadIdx.x < 16) db_s[threadIdx.x] = 0.f;
    if (warp == 1) { tc::tmem_alloc(&tmem_s, 128); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::mbar_expect_tx(&wbar, DG_W_BYTES);
            tc::tma_load_2d(smem_w, &map_w, &wbar, 0, 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int stage = it % DG_STAGES, phase = (it / DG_STAGES) & 1;
                tc::mbar_wait(&empty[stage], phase ^ 1);
                tc::mbar_expect_tx(&full[stage], DG_A_BYTES);
                tc::tma_load_2d(smem_a + stage * DG_A_BYTES, &map_dz2, &full[stage], 0, tile * 128 - 14);   // rows < 0 are zero-filled
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint64_t KM64 = tc::smem_desc_base(0, 512, tc::SW_64B);
            constexpr uint32_t IDESC = tc::idesc_bf16(128, 64);
            const uint32_t w_addr = tc::smem_u32(smem_w), a_addr = tc::smem_u32(smem_a);
            tc::mbar_wait(&wbar, 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int stage = it % DG_STAGES, phase = (it / DG_STAGES) & 1;
                const int acc = it & 1, acc_phase = (it >> 1) & 1;
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::mbar_wait(&full[stage], phase);
                tc::tc_fence_after();
                const uint32_t a_base = a_addr + stage * DG_A_BYTES;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int shift = 14 - ((t >> 1) * SG_W + (t & 1));        // input row R reads DZ2 row R - (ty*13+tx)
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc::mma_f16_ss(tmem + acc * 64, tc::smem_desc(KM64, a_base + shift * 64 + k * 32),
                                       tc::smem_desc(KM64, w_addr + t * 4096 + k * 32), IDESC, (t | k) != 0);
                }
                tc::mma_commit(&empty[stage]);
                tc::mma_commit(&tfull[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        float dbacc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) dbacc[c] = 0.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1, acc_phase = (it >> 1) & 1;
            const int row = tile * 128 + q * 32 + lane;
            const int b = row / ROWS_PER_SAMPLE, rr = row - b * ROWS_PER_SAMPLE;
            const int sy = rr / SG_W, sx = rr - sy * SG_W;
            const bool in = row < total_rows;
            tc::mbar_wait(&tfull[acc], acc_phase);
            tc::tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {                   // half = dy: cells (dy,0) and (dy,1), 16 channels each
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + acc * 64 + half * 32, r);
                tc::tmem_ld_wait();
                if (half == 1) {
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&tempty[acc]);
                }
                if (in) {
                    const uint4* ap = reinterpret_cast<const uint4*>(p1 + (size_t)row * 64 + half * 32);
                    const int oy = 2 * sy + half - 1;                 // cell (dy = half, dx) of super pixel (sy,sx) is conv1 output (oy, ox)
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        const int ox = 2 * sx + dx - 1;
                        if (oy < 0 || oy >= 64 || ox < 0 || ox >= 25) continue;      // zero-padding cells of P1: no conv1 output behind them
                        uint4* op = reinterpret_cast<uint4*>(dz1 + ((size_t)b * 1690 + oy * 26 + ox) * 16);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {                 // 8 channels per 16-byte vector
                            const uint4 av = ap[dx * 2 + h];
                            const uint32_t au[4] = {av.x, av.y, av.z, av.w};
                            uint32_t o[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = dx * 16 + h * 8 + 2 * j;
                                const float g0 = __uint_as_float(r[c]) * (bf_lo(au[j]) > 0.f ? 1.f : 0.2f);
                                const float g1 = __uint_as_float(r[c + 1]) * (bf_hi(au[j]) > 0.f ? 1.f : 0.2f);
                                o[j] = pack_bf16x2(g0, g1);
                                dbacc[h * 8 + 2 * j] += bf_lo(o[j]);
                                dbacc[h * 8 + 2 * j + 1] += bf_hi(o[j]);
                            }
                            op[h] = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float s = warp_sum(dbacc[c]);
            if (lane == 0) atomicAdd(&db_s[c], s);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 128);
    if (threadIdx.x < 16) atomicAdd(&db1[threadIdx.x], db_s[threadIdx.x]);
}

}  // namespace

extern "C" {

// a2 (B*429,32) bf16, dlogit (B,) fp32 -> dz2 (B*429,32) bf16;  dfc_w (1,12288) += ;  dconv2_b (32,) +=   (fp32, caller zeroes or accumulates)
int mmg_disc_fc_bwd(const void* a2, const float* dlogit, const void* packed, void* dz2, float* dwfcp, float* dconv2_b, int64_t B, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMG_REQUIRE(a2 && dlogit && packed && dz2 && dwfcp && dconv2_b && B >= 0, MMG_EINVAL, "fc_bwd: bad arguments");
    if (B == 0) return MMG_OK;
    const float* wfcp = (const float*)((const unsigned char*)packed + 2048 + 32768);
    const int gx = (ROWS_PER_SAMPLE + 63) / 64;
    int gy = (4 * MMG_NUM_SMS + gx - 1) / gx;
    if (gy > B) gy = (int)B;
    const int bpb = (int)((B + gy - 1) / gy);
    gy = (int)((B + bpb - 1) / bpb);
    fc_bwd_kernel<<<dim3(gx, gy), 256, 0, stream>>>((const __nv_bfloat16*)a2, dlogit, wfcp, (__nv_bfloat16*)dz2, dwfcp, dconv2_b, (int)B, bpb);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// dconv2_w (32,16,4,4) fp32 += sum over rows  (caller zeroes or accumulates)
int mmg_disc_conv2_wgrad(const void* p1, const void* dz2, float* dconv2_w, int64_t B, void* stream) {
    MMG_REQUIRE(p1 && dz2 && dconv2_w && B >= 0, MMG_EINVAL, "conv2_wgrad: bad arguments");
    if (B == 0) return MMG_OK;
    const int64_t rows = B * ROWS_PER_SAMPLE;
    MMG_REQUIRE(rows < (1LL << 31) - 256, MMG_EUNSUPPORTED, "conv2_wgrad: batch too large");
    CUtensorMap map_p1, map_dz2;
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_p1, p1, 64, (uint64_t)rows, 128, 64, 144, CU_TENSOR_MAP_SWIZZLE_128B) == 0, MMG_EINVAL, "conv2_wgrad: tensor map (p1)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_dz2, dz2, 32, (uint64_t)rows, 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B) == 0, MMG_EINVAL, "conv2_wgrad: tensor map (dz2)");
    const int chunks = (int)((rows + 127) / 128);
    const int grid = chunks < MMG_NUM_SMS ? chunks : MMG_NUM_SMS;
    MMG_CUDA(cudaFuncSetAttribute(conv2_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    conv2_wgrad_tc_kernel<<<grid, 192, WG_SMEM, (cudaStream_t)stream>>>(map_p1, map_dz2, dconv2_w, chunks);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// dz2 (B*429,32), p1 (B*429,64) -> dz1c (B*1690,16) bf16 in conv1's row space (junk rows untouched: allocate zeroed);  dconv1_b (16,) fp32 +=
int mmg_disc_conv2_dgrad(const void* dz2, const void* packed, const void* p1, void* dz1, float* dconv1_b, int64_t B, void* stream) {
    MMG_REQUIRE(dz2 && packed && p1 && dz1 && dconv1_b && B >= 0, MMG_EINVAL, "conv2_dgrad: bad arguments");
    if (B == 0) return MMG_OK;
    const int64_t rows = B * ROWS_PER_SAMPLE;
    MMG_REQUIRE(rows < (1LL << 31) - 256, MMG_EUNSUPPORTED, "conv2_dgrad: batch too large");
    const unsigned char* pk = (const unsigned char*)packed;
    CUtensorMap map_dz2, map_w;
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_dz2, dz2, 32, (uint64_t)rows, 64, 32, 144, CU_TENSOR_MAP_SWIZZLE_64B) == 0, MMG_EINVAL, "conv2_dgrad: tensor map (dz2)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w, pk + 2048 + 16384, 32, 256, 64, 32, 256, CU_TENSOR_MAP_SWIZZLE_64B) == 0, MMG_EINVAL, "conv2_dgrad: tensor map (w2d)");
    const int tiles = (int)((rows + 127) / 128);
    const int grid = tiles < 2 * MMG_NUM_SMS ? tiles : 2 * MMG_NUM_SMS;
    MMG_CUDA(cudaFuncSetAttribute(conv2_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM));
    conv2_dgrad_tc_kernel<<<grid, 192, DG_SMEM, (cudaStream_t)stream>>>(map_dz2, map_w, (const __nv_bfloat16*)p1, (__nv_bfloat16*)dz1, dconv1_b, (int)rows, tiles);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
