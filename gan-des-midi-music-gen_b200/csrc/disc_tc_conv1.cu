// conv1 of DiscriminatorCNN (Conv2d 2->16, k4 s2 p1, network_tests.py:150,157) on the tcgen05 tensor cores.
//
//   xs_pack      (elementwise)  X (B,2,128,50) u8|f32  ->  XS (B*1690, 8) bf16: zero-padded input as 65x26 super pixels
//   conv1 fwd    (tcgen05)      row m = b*1690 + oy*26 + ox;  D[m][oc] = sum_{ty} XS[m + 26*ty .. +1][16] . W1b[ty][oc][16]
//                               A: un-swizzled K-major, 16-byte rows, second K chunk = next row (LBO = 16 B, SBO = 128 B)
//                               B: SW32 K-major.  Epilogue: bias + LeakyReLU -> bf16 -> P1 (the layout conv2 consumes)
//   conv1 wgrad  (tcgen05)      dW1[k][oc] = sum_m XS[m + ...][k] * DZ1c[m][oc]: both operands MN-major; A is the same XS box
//                               (M = 64 = 8 atoms one row apart, of which atoms 0/1 = the two horizontal taps are used),
//                               B = DZ1c rows (SW32).  Accumulators stay in TMEM for the CTA's whole slice of the batch.
// Operand layouts validated on hardware by tools/tc_probe.cu (experiments E7b, E8).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int XS_ROWS = 1690;            // 65 x 26 super pixels per sample
constexpr int XS_W = 26;
constexpr int P1_ROWS = 429;
constexpr int P1_W = 13;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// one thread per XS row: 8 values (dy,dx,ch) of super pixel (sy,sx); y' = 2sy+dy = iy+1, x' = 2sx+dx = ix+1
template <typename InT>
__global__ void __launch_bounds__(256) xs_pack_kernel(const InT* __restrict__ x, __nv_bfloat16* __restrict__ xs, long long total_rows) {
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < total_rows; row += (long long)gridDim.x * blockDim.x) {
        const long long b = row / XS_ROWS;
        const int rr = (int)(row - b * XS_ROWS), sy = rr / XS_W, sx = rr - sy * XS_W;
        const InT* xb = x + (size_t)b * 2 * 128 * 50;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int dy = e >> 2, dx = (e >> 1) & 1, ch = e & 1;
            const int iy = 2 * sy + dy - 1, ix = 2 * sx + dx - 1;
            v[e] = (iy >= 0 && iy < 128 && ix >= 0 && ix < 50) ? (float)xb[(ch * 128 + iy) * 50 + ix] : 0.f;
        }
        *reinterpret_cast<uint4*>(xs + row * 8) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int C1_STAGES = 8;
constexpr int C1_BOX_ROWS = 160;                         // 128 + 27 halo, padded
constexpr int C1_A_BYTES = C1_BOX_ROWS * 16;
constexpr int C1F_SMEM = 1024 + C1_STAGES * C1_A_BYTES + 1024;

__global__ void __launch_bounds__(192, 2) conv1_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_xs, const __grid_constant__ CUtensorMap map_w,
                                                              const float* __restrict__ bias, __nv_bfloat16* __restrict__ p1, long long total_rows,
                                                              int num_tiles) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full[C1_STAGES], empty[C1_STAGES], tfull[2], tempty[2], wbar;
    __shared__ uint32_t tmem_s;
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* smem_w = smem;                          // 1 KB: [2 ty][16 oc][16 k] bf16, SW32
    unsigned char* smem_a = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < C1_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
        tc::mbar_init(&wbar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_s, 32); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::mbar_expect_tx(&wbar, 1024);
            tc::tma_load_2d(smem_w, &map_w, &wbar, 0, 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int stage = it % C1_STAGES, phase = (it / C1_STAGES) & 1;
                tc::mbar_wait(&empty[stage], phase ^ 1);
                tc::mbar_expect_tx(&full[stage], C1_A_BYTES);
                tc::tma_load_2d(smem_a + stage * C1_A_BYTES, &map_xs, &full[stage], 0, tile * 128);
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint64_t A_K = tc::smem_desc_base(16, 128, tc::SW_NONE);     // K chunk 1 = the next 16-byte row
            constexpr uint64_t B_K = tc::smem_desc_base(0, 256, tc::SW_32B);
            constexpr uint32_t IDESC = tc::idesc_bf16(128, 16);
            const uint32_t w_addr = tc::smem_u32(smem_w), a_addr = tc::smem_u32(smem_a);
            tc::mbar_wait(&wbar, 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int stage = it % C1_STAGES, phase = (it / C1_STAGES) & 1;
                const int acc = it & 1, acc_phase = (it >> 1) & 1;
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::mbar_wait(&full[stage], phase);
                tc::tc_fence_after();
                const uint32_t a_base = a_addr + stage * C1_A_BYTES;
#pragma unroll
                for (int ty = 0; ty < 2; ++ty)
                    tc::mma_f16_ss(tmem + acc * 16, tc::smem_desc(A_K, a_base + ty * XS_W * 16), tc::smem_desc(B_K, w_addr + ty * 512), IDESC, ty != 0);
                tc::mma_commit(&empty[stage]);
                tc::mma_commit(&tfull[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        float bs[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) bs[c] = bias[c];
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1, acc_phase = (it >> 1) & 1;
            tc::mbar_wait(&tfull[acc], acc_phase);
            tc::tc_fence_after();
            uint32_t r[16];
            tc::tmem_ld_32x16(tmem + ((uint32_t)(q * 32) << 16) + acc * 16, r);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
            const long long row = (long long)tile * 128 + q * 32 + lane;
            const long long b = row / XS_ROWS;
            const int rr = (int)(row - b * XS_ROWS), oy = rr / XS_W, ox = rr - oy * XS_W;
            if (row < total_rows && oy < 64 && ox < 25) {
                uint32_t o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float z0 = __uint_as_float(r[2 * c]) + bs[2 * c], z1 = __uint_as_float(r[2 * c + 1]) + bs[2 * c + 1];
                    o[c] = pack_bf16x2(z0 > 0.f ? z0 : 0.2f * z0, z1 > 0.f ? z1 : 0.2f * z1);
                }
                const int yp = oy + 1, xp = ox + 1;
                const size_t prow = (size_t)b * P1_ROWS + (yp >> 1) * P1_W + (xp >> 1);
                uint4* dst = reinterpret_cast<uint4*>(p1 + prow * 64 + ((yp & 1) * 2 + (xp & 1)) * 16);
                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 32);
}

// ------------------------------------------------------------------------------------------------
constexpr int C1W_B_BYTES = 128 * 32;                    // DZ1c box
constexpr int C1W_STAGE_BYTES = C1W_B_BYTES + C1_A_BYTES;  // 6656 = 26 * 256
constexpr int C1W_SMEM = 1024 + C1_STAGES * C1W_STAGE_BYTES + 256;

__global__ void __launch_bounds__(192, 2) conv1_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_xs, const __grid_constant__ CUtensorMap map_dz,
                                                                float* __restrict__ dw1, int num_chunks) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full[C1_STAGES], empty[C1_STAGES], done;
    __shared__ uint32_t tmem_s;
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c_lo = (int)((long long)num_chunks * blockIdx.x / gridDim.x), c_hi = (int)((long long)num_chunks * (blockIdx.x + 1) / gridDim.x);
    if (threadIdx.x == 0) {
        for (int i = 0; i < C1_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_s, 32); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;
    if (c_hi > c_lo) {
        if (warp == 0 && tc::elect_one()) {
            for (int c = c_lo, it = 0; c < c_hi; ++c, ++it) {
                const int stage = it % C1_STAGES, phase = (it / C1_STAGES) & 1;
                tc::mbar_wait(&empty[stage], phase ^ 1);
                tc::mbar_expect_tx(&full[stage], C1W_STAGE_BYTES);
                unsigned char* st = smem + stage * C1W_STAGE_BYTES;
                tc::tma_load_2d(st, &map_dz, &full[stage], 0, c * 128);
                tc::tma_load_2d(st + C1W_B_BYTES, &map_xs, &full[stage], 0, c * 128);
            }
        } else if (warp == 1 && tc::elect_one()) {
            constexpr uint64_t A_MN = tc::smem_desc_base(128, 16, tc::SW_NONE);    // atoms one row (16 B) apart, 8-row K groups 128 B apart
            constexpr uint64_t B_MN = tc::smem_desc_base(0, 256, tc::SW_32B);
            constexpr uint32_t IDESC = tc::idesc_bf16(64, 16, 1, 1);
            for (int c = c_lo, it = 0; c < c_hi; ++c, ++it) {
                const int stage = it % C1_STAGES, phase = (it / C1_STAGES) & 1;
                tc::mbar_wait(&full[stage], phase);
                tc::tc_fence_after();
                const uint32_t b_base = tc::smem_u32(smem + stage * C1W_STAGE_BYTES), a_base = b_base + C1W_B_BYTES;
#pragma unroll
                for (int ty = 0; ty < 2; ++ty)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        tc::mma_f16_ss(tmem + ty * 16, tc::smem_desc(A_MN, a_base + (ty * XS_W + k * 16) * 16), tc::smem_desc(B_MN, b_base + k * 16 * 32), IDESC,
                                       (it | k) != 0);
                tc::mma_commit(&empty[stage]);
            }
            tc::mma_commit(&done);
        } else if (warp >= 2 && (warp & 3) == 0) {
            // M = 64 accumulator: row i sits in TMEM lane (i % 16) + 32 * (i / 16); rows 0..15 (atoms 0,1 = tx) are the useful ones
            tc::mbar_wait(&done, 0);
            tc::tc_fence_after();
#pragma unroll
            for (int ty = 0; ty < 2; ++ty) {
                uint32_t r[16];
                tc::tmem_ld_32x16(tmem + ty * 16, r);
                tc::tmem_ld_wait();
                if (lane < 16) {
                    const int tx = lane >> 3, e = lane & 7, dy = e >> 2, dx = (e >> 1) & 1, ch = e & 1;
#pragma unroll
                    for (int oc = 0; oc < 16; ++oc)
                        atomicAdd(&dw1[((oc * 2 + ch) * 4 + 2 * ty + dy) * 4 + 2 * tx + dx], __uint_as_float(r[oc]));
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 32);
}

int make_xs_map(CUtensorMap* map, const void* xs, uint64_t rows) {
    auto fn = tc::get_encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[2] = {8, rows};
    cuuint64_t strides[1] = {16};
    cuuint32_t box[2] = {8, C1_BOX_ROWS};
    cuuint32_t es[2] = {1, 1};
    return (int)fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(xs), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

extern "C" {

// x: (B,2,128,50) uint8 (x_dtype 2) or float32 (x_dtype 0)  ->  xs (B*1690, 8) bf16
int mmg_disc_xs_pack(const void* x, int x_dtype, void* xs, int64_t B, void* stream) {
    MMG_REQUIRE(x && xs && B >= 0, MMG_EINVAL, "xs_pack: bad arguments");
    if (B == 0) return MMG_OK;
    const long long rows = B * XS_ROWS;
    if (x_dtype == 2)
        xs_pack_kernel<uint8_t><<<mmg_grid(rows, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)x, (__nv_bfloat16*)xs, rows);
    else if (x_dtype == 0)
        xs_pack_kernel<float><<<mmg_grid(rows, 256), 256, 0, (cudaStream_t)stream>>>((const float*)x, (__nv_bfloat16*)xs, rows);
    else
        MMG_REQUIRE(false, MMG_EINVAL, "xs_pack: x_dtype must be 0 (f32) or 2 (u8)");
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// xs (B*1690,8) bf16 -> p1 (B*429,64) bf16 = LeakyReLU(conv1(x)) in the space-to-depth layout (pad cells of p1 must already be zero)
int mmg_disc_conv1_fwd(const void* xs, const void* packed, const float* conv1_b, void* p1, int64_t B, void* stream) {
    MMG_REQUIRE(xs && packed && conv1_b && p1 && B >= 0, MMG_EINVAL, "conv1_fwd: bad arguments");
    if (B == 0) return MMG_OK;
    const long long rows = B * XS_ROWS;
    MMG_REQUIRE(rows < (1LL << 31) - 256, MMG_EUNSUPPORTED, "conv1_fwd: batch too large");
    CUtensorMap map_xs, map_w;
    MMG_REQUIRE(make_xs_map(&map_xs, xs, (uint64_t)rows) == 0, MMG_EINVAL, "conv1_fwd: tensor map (xs)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w, packed, 16, 32, 32, 16, 32, CU_TENSOR_MAP_SWIZZLE_32B) == 0, MMG_EINVAL, "conv1_fwd: tensor map (w1b)");
    const int tiles = (int)((rows + 127) / 128);
    const int grid = tiles < 2 * MMG_NUM_SMS ? tiles : 2 * MMG_NUM_SMS;
    MMG_CUDA(cudaFuncSetAttribute(conv1_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C1F_SMEM));
    conv1_fwd_tc_kernel<<<grid, 192, C1F_SMEM, (cudaStream_t)stream>>>(map_xs, map_w, conv1_b, (__nv_bfloat16*)p1, rows, tiles);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// xs (B*1690,8), dz1c (B*1690,16) bf16 (junk rows zero) -> dconv1_w (16,2,4,4) fp32 +=
int mmg_disc_conv1_wgrad(const void* xs, const void* dz1c, float* dconv1_w, int64_t B, void* stream) {
    MMG_REQUIRE(xs && dz1c && dconv1_w && B >= 0, MMG_EINVAL, "conv1_wgrad: bad arguments");
    if (B == 0) return MMG_OK;
    const long long rows = B * XS_ROWS;
    MMG_REQUIRE(rows < (1LL << 31) - 256, MMG_EUNSUPPORTED, "conv1_wgrad: batch too large");
    CUtensorMap map_xs, map_dz;
    MMG_REQUIRE(make_xs_map(&map_xs, xs, (uint64_t)rows) == 0, MMG_EINVAL, "conv1_wgrad: tensor map (xs)");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_dz, dz1c, 16, (uint64_t)rows, 32, 16, 128, CU_TENSOR_MAP_SWIZZLE_32B) == 0, MMG_EINVAL, "conv1_wgrad: tensor map (dz1c)");
    const int chunks = (int)((rows + 127) / 128);
    const int grid = chunks < 2 * MMG_NUM_SMS ? chunks : 2 * MMG_NUM_SMS;
    MMG_CUDA(cudaFuncSetAttribute(conv1_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C1W_SMEM));
    conv1_wgrad_tc_kernel<<<grid, 192, C1W_SMEM, (cudaStream_t)stream>>>(map_xs, map_dz, dconv1_w, chunks);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
