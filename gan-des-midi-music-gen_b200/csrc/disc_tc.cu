// bf16 tensor-core path of the MM-GAN discriminator (DiscriminatorCNN, network_tests.py:147-160):
//   conv1 2->16 k4 s2 p1 + LeakyReLU   (tcgen05 tap-shift GEMM over the space-to-depth input XS, K = 2 x 16)
//   conv2 16->32 k4 s2 p1 + LeakyReLU  (tcgen05 "tap-shift" implicit GEMM fed by TMA) + fc partial dot
//
// Data layout in HBM (all bf16 unless noted) -- "S2D" = padded space-to-depth(2):
//   X    (B,2,128,50) uint8          piano roll / duration planes (values are small integers, exact in bf16)
//   XS   (B*1690, 8)                 the zero-padded input (y' = iy+1, x' = ix+1) as 65x26 super pixels of 2x2 px x 2 ch:
//                                    8 values = (dy,dx,ch), 16 bytes per row.  conv1 output row m = b*1690 + oy*26 + ox reads rows
//                                    m + {0,1,26,27}; the two horizontally adjacent taps are 32 contiguous bytes = one K=16 step,
//                                    expressed as an un-swizzled K-major operand whose second K chunk starts 16 B later (LBO=16).
//   P1   (B*429, 64)                 conv1 activations.  Row R = b*429 + sy*13 + sx is one 2x2 super pixel of the
//                                    zero-padded 66x26 map (y' = iy+1, x' = ix+1), 64 values = (dy,dx,c16).
//                                    A stride-2 4x4 conv over the 64x25 map is then a stride-1 2x2 conv over the
//                                    33x13 super-pixel grid with 64 channels: output row m = b*429 + oy*13 + ox
//                                    reads super pixels m + {0,1,13,14}  ->  four K=64 GEMM taps whose A operand
//                                    is the SAME shared-memory box, addressed with a row-shifted UMMA descriptor.
//   A2   (B*429, 32)                 conv2 activations in the same row space (ox==12 / oy==32 rows are junk = 0).
//   logits (B,) fp32                 fc output without bias, accumulated with atomics from the conv2 epilogue.
// Weights are repacked from the fp32 master tensors (the nn.Parameters) by mmg_disc_pack_weights.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int ROWS_PER_SAMPLE = 429;      // 33 x 13 super pixels
constexpr int SG_W = 13;

// ------------------------------------------------------------------------------------------------
// weight packing (fp32 master -> operand layouts)
// ------------------------------------------------------------------------------------------------
// w1b  bf16 [2 ty][16 oc][16 k]  k = tx*8 + (dy*2+dx)*2 + ch for ky = 2ty+dy, kx = 2tx+dx                          (conv1 fwd B)
// w2p  bf16 [4 t][32 oc][64 k]   t = (ky>>1)*2 + (kx>>1), k = ((ky&1)*2 + (kx&1))*16 + ic      (conv2 fwd B, wgrad layout)
// w2d  bf16 [4 t][64 n][32 oc]   n = ((ky&1)*2 + (kx&1))*16 + ic                                  (conv2 dgrad B)
// wfcp fp32 [429 r][32 oc]       fc.weight[oc*384 + oy*12 + ox] at r = oy*13 + ox, 0 on junk rows
__global__ void pack_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ wfc,
                                    __nv_bfloat16* __restrict__ w1b, __nv_bfloat16* __restrict__ w2p, __nv_bfloat16* __restrict__ w2d,
                                    float* __restrict__ wfcp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 512) {                                       // conv1.weight (16,2,4,4)
        const int oc = i / 32, ch = (i / 16) % 2, ky = (i / 4) % 4, kx = i % 4;
        w1b[((ky >> 1) * 16 + oc) * 16 + (kx >> 1) * 8 + ((ky & 1) * 2 + (kx & 1)) * 2 + ch] = __float2bfloat16(w1[i]);
    }
    if (i < 8192) {                                      // conv2.weight (32,16,4,4)
        const int oc = i / 256, ic = (i / 16) % 16, ky = (i / 4) % 4, kx = i % 4;
        const int t = (ky >> 1) * 2 + (kx >> 1), k = ((ky & 1) * 2 + (kx & 1)) * 16 + ic;
        const __nv_bfloat16 v = __float2bfloat16(w2[i]);
        w2p[(t * 32 + oc) * 64 + k] = v;
        w2d[(t * 64 + k) * 32 + oc] = v;
    }
    if (i < ROWS_PER_SAMPLE * 32) {
        const int r = i / 32, oc = i % 32, oy = r / SG_W, ox = r % SG_W;
        wfcp[i] = (oy < 32 && ox < 12) ? wfc[oc * 384 + oy * 12 + ox] : 0.f;
    }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------
// conv2 forward: tcgen05 tap-shift GEMM.  M = 128 output rows per tile, N = 32, K = 4 taps x 64.
// warp 0: TMA producer, warp 1: MMA issuer (+TMEM owner), warps 2-5: epilogue (one TMEM lane quadrant each)
// ------------------------------------------------------------------------------------------------
constexpr int C2F_STAGES = 4;
constexpr int C2F_BOX_ROWS = 144;                       // 128 + 14 halo rows, padded to a multiple of 8
constexpr int C2F_STAGE_BYTES = C2F_BOX_ROWS * 128;
constexpr int C2F_W_BYTES = 4 * 32 * 128;
constexpr int C2F_SMEM = C2F_W_BYTES + C2F_STAGES * C2F_STAGE_BYTES + 1024;

__global__ void __launch_bounds__(192, 2) conv2_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_p1, const __grid_constant__ CUtensorMap map_w,
                                                              const float* __restrict__ bias, const float* __restrict__ wfcp,
                                                              __nv_bfloat16* __restrict__ a2, float* __restrict__ logits, int total_rows, int num_tiles) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full[C2F_STAGES], empty[C2F_STAGES], tfull[2], tempty[2], wbar;
    __shared__ uint32_t tmem_s;
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* smem_w = smem;
    unsigned char* smem_a = smem + C2F_W_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C2F_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
        tc::mbar_init(&wbar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_s, 64); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&map_p1);
            tc::mbar_expect_tx(&wbar, C2F_W_BYTES);
            tc::tma_load_2d(smem_w, &map_w, &wbar, 0, 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int stage = it % C2F_STAGES, phase = (it / C2F_STAGES) & 1;
                tc::mbar_wait(&empty[stage], phase ^ 1);
                tc::mbar_expect_tx(&full[stage], C2F_STAGE_BYTES);
                tc::tma_load_2d(smem_a + stage * C2F_STAGE_BYTES, &map_p1, &full[stage], 0, tile * 128);
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint64_t KM128 = tc::smem_desc_base(0, 1024, tc::SW_128B);
            constexpr uint32_t IDESC = tc::idesc_bf16(128, 32);
            const uint32_t w_addr = tc::smem_u32(smem_w), a_addr = tc::smem_u32(smem_a);
            tc::mbar_wait(&wbar, 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int stage = it % C2F_STAGES, phase = (it / C2F_STAGES) & 1;
                const int acc = it & 1, acc_phase = (it >> 1) & 1;
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::mbar_wait(&full[stage], phase);
                tc::tc_fence_after();
                const uint32_t a_base = a_addr + stage * C2F_STAGE_BYTES;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int shift = (t >> 1) * SG_W + (t & 1);          // super-pixel row offset of tap (ty,tx)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::mma_f16_ss(tmem + acc * 32, tc::smem_desc(KM128, a_base + shift * 128 + k * 32),
                                       tc::smem_desc(KM128, w_addr + t * 4096 + k * 32), IDESC, (t | k) != 0);
                }
                tc::mma_commit(&empty[stage]);
                tc::mma_commit(&tfull[acc]);
            }
        }
    } else {
        const int q = warp & 3;                            // TMEM lane quadrant this warp may read
        float bs[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) bs[c] = bias[c];
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1, acc_phase = (it >> 1) & 1;
            tc::mbar_wait(&tfull[acc], acc_phase);
            tc::tc_fence_after();
            uint32_t r[32];
            tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + acc * 32, r);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
            const int row = tile * 128 + q * 32 + lane;
            const int b = row / ROWS_PER_SAMPLE, rr = row - b * ROWS_PER_SAMPLE;
            const int oy = rr / SG_W, ox = rr - oy * SG_W;
            const bool real = row < total_rows && oy < 32 && ox < 12;
            float dot = 0.f;
            uint32_t o[16];
            const float4* wf = reinterpret_cast<const float4*>(wfcp + (size_t)rr * 32);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 w = real ? wf[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float z = __uint_as_float(r[4 * c4 + j]) + bs[4 * c4 + j];
                    z = z > 0.f ? z : 0.2f * z;
                    v[j] = real ? __bfloat162float(__float2bfloat16(z)) : 0.f;     // what the backward will read
                }
                dot = fmaf(v[0], w.x, dot); dot = fmaf(v[1], w.y, dot); dot = fmaf(v[2], w.z, dot); dot = fmaf(v[3], w.w, dot);
                o[2 * c4] = pack_bf16x2(v[0], v[1]);
                o[2 * c4 + 1] = pack_bf16x2(v[2], v[3]);
            }
            if (row < total_rows) {
                uint4* dst = reinterpret_cast<uint4*>(a2 + (size_t)row * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
            // per-sample reduction of the fc partial dot: a warp's 32 rows span at most two samples
            const int b0 = __shfl_sync(0xffffffffu, b, 0);
            const float s0 = warp_sum(b == b0 ? dot : 0.f), s1 = warp_sum(b != b0 ? dot : 0.f);
            if (lane == 0) {
                if ((size_t)b0 * ROWS_PER_SAMPLE < (size_t)total_rows) atomicAdd(&logits[b0], s0);
                if (s1 != 0.f) atomicAdd(&logits[b0 + 1], s1);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 64);
}

}  // namespace

extern "C" {

size_t mmg_disc_packed_weights_bytes(void) { return 512 * 4 + 8192 * 2 + 8192 * 2 + ROWS_PER_SAMPLE * 32 * 4; }

// packed: [w1b bf16 512, padded to 2 KB][w2p bf16 8192][w2d bf16 8192][wfcp fp32 429*32]
int mmg_disc_pack_weights(const float* conv1_w, const float* conv2_w, const float* fc_w, void* packed, void* stream) {
    MMG_REQUIRE(conv1_w && conv2_w && fc_w && packed, MMG_EINVAL, "pack_weights: null pointer");
    unsigned char* p = (unsigned char*)packed;
    pack_weights_kernel<<<(ROWS_PER_SAMPLE * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        conv1_w, conv2_w, fc_w, (__nv_bfloat16*)p, (__nv_bfloat16*)(p + 2048), (__nv_bfloat16*)(p + 2048 + 16384), (float*)(p + 2048 + 32768));
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

// p1 (B*429,64) bf16 -> a2 (B*429,32) bf16 (junk rows zero) and logits[b] += <a2_b, fc.weight>  (logits must be pre-zeroed)
int mmg_disc_conv2_fwd(const void* p1, const void* packed, const float* conv2_b, void* a2, float* logits, int64_t B, void* stream) {
    MMG_REQUIRE(p1 && packed && conv2_b && a2 && logits && B >= 0, MMG_EINVAL, "conv2_fwd: bad arguments");
    if (B == 0) return MMG_OK;
    const int64_t rows = B * ROWS_PER_SAMPLE;
    MMG_REQUIRE(rows < (1LL << 31) - 256, MMG_EUNSUPPORTED, "conv2_fwd: batch too large");
    const unsigned char* pk = (const unsigned char*)packed;
    CUtensorMap map_p1, map_w;
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_p1, p1, 64, (uint64_t)rows, 128, 64, C2F_BOX_ROWS, CU_TENSOR_MAP_SWIZZLE_128B) == 0, MMG_EINVAL,
                "conv2_fwd: cuTensorMapEncodeTiled(p1) failed");
    MMG_REQUIRE(tc::make_map_2d_bf16(&map_w, pk + 2048, 64, 128, 128, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B) == 0, MMG_EINVAL,
                "conv2_fwd: cuTensorMapEncodeTiled(w2p) failed");
    const int tiles = (int)((rows + 127) / 128);
    const int grid = tiles < 2 * MMG_NUM_SMS ? tiles : 2 * MMG_NUM_SMS;
    MMG_CUDA(cudaFuncSetAttribute(conv2_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C2F_SMEM));
    conv2_fwd_tc_kernel<<<grid, 192, C2F_SMEM, (cudaStream_t)stream>>>(map_p1, map_w, conv2_b, (const float*)(pk + 2048 + 32768), (__nv_bfloat16*)a2, logits,
                                                                       (int)rows, tiles);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
