// bf16 tensor-core path of the MM-GAN discriminator, backward (autograd of network_tests.py:156-160,
// reached by disc_loss.backward() / gen_loss.backward(), network_tests.py:307,314).  Layouts: see disc_tc.cu.
//
//   fc_bwd      (SIMT, HBM-bound)  dz2 = dlogit[b] * fc.w * lrelu'(a2)  -> DZ2 bf16;  dfc.w += dlogit[b]*a2;  dconv2.b += dz2
//   conv2 wgrad (tcgen05)          dW2[t][k64][oc] = sum_rows P1[row+shift_t][k] * DZ2[row][oc]    both operands MN-major:
//                                  M = 128 = the two horizontally adjacent taps (second atom = same box, one row later)
//   conv2 dgrad (tcgen05)          da1[R][n64] = sum_t DZ2[R-shift_t][oc] * W2d[t][n][oc];  epilogue: dz1 = da1*lrelu'(a1),
//                                  pad cells -> 0, dconv1.b += dz1, DZ1 bf16 written in the P1 layout
//   conv1 wgrad (SIMT)             dW1[k32][oc16] = sum_pos x_patch[pos][k] * dz1[pos][oc]
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
        if (warp == 0 && lane == 0) {
            for (int c = c_lo, it = 0; c < c_hi; ++c, ++it) {
                const int stage = it % WG_STAGES, phase = (it / WG_STAGES) & 1;
                tc::mbar_wait(&empty[stage], phase ^ 1);
                tc::mbar_expect_tx(&full[stage], WG_STAGE_BYTES);
                unsigned char* st = smem + stage * WG_STAGE_BYTES;
                tc::tma_load_2d(st, &map_p1, &full[stage], 0, c * 128);
                tc::tma_load_2d(st + WG_A_BYTES, &map_dz2, &full[stage], 0, c * 128);
            }
        } else if (warp == 1 && lane == 0) {
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
        if (lane == 0) {
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
        if (lane == 0) {
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
                    uint4* op = reinterpret_cast<uint4*>(dz1 + (size_t)row * 64 + half * 32);
                    const bool pad_y = (sy == 0 && half == 0) || (sy == 32 && half == 1);
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        const bool pad = pad_y || (sx == 0 && dx == 0);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {                 // 8 channels per 16-byte vector
                            const uint4 av = ap[dx * 2 + h];
                            const uint32_t au[4] = {av.x, av.y, av.z, av.w};
                            uint32_t o[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = dx * 16 + h * 8 + 2 * j;
                                float g0 = __uint_as_float(r[c]) * (bf_lo(au[j]) > 0.f ? 1.f : 0.2f);
                                float g1 = __uint_as_float(r[c + 1]) * (bf_hi(au[j]) > 0.f ? 1.f : 0.2f);
                                if (pad) { g0 = 0.f; g1 = 0.f; }
                                o[j] = pack_bf16x2(g0, g1);
                                dbacc[h * 8 + 2 * j] += bf_lo(o[j]);
                                dbacc[h * 8 + 2 * j + 1] += bf_hi(o[j]);
                            }
                            op[dx * 2 + h] = make_uint4(o[0], o[1], o[2], o[3]);
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

// ------------------------------------------------------------------------------------------------
// conv1 weight gradient (SIMT): lane = patch index k (32), 16 channel accumulators per lane, a warp walks
// output positions; DZ1 of the sample is staged in shared memory as fp32, the input planes as bf16.
// ------------------------------------------------------------------------------------------------
constexpr int C1W_THREADS = 512;
constexpr int C1W_SMEM = 1600 * 16 * 4 + 2 * 130 * 52 * 2;

template <typename InT>
__global__ void __launch_bounds__(C1W_THREADS, 1) conv1_wgrad_kernel(const InT* __restrict__ x, const __nv_bfloat16* __restrict__ dz1,
                                                                      float* __restrict__ dw1, int B) {
    extern __shared__ __align__(16) unsigned char sm[];
    float* gs = reinterpret_cast<float*>(sm);                                   // [1600 pos][16 oc]
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(sm + 1600 * 16 * 4);   // [2][130][52]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = lane >> 4, ky = (lane >> 2) & 3, kx = lane & 3;              // k = (ch*4+ky)*4+kx
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    for (int i = tid; i < 2 * 130 * 52; i += C1W_THREADS) xs[i] = __float2bfloat16(0.f);
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        const InT* xb = x + (size_t)b * 2 * 128 * 50;
        for (int i = tid; i < 2 * 128 * 50; i += C1W_THREADS) {
            const int c = i / 6400, iy = (i / 50) % 128, ix = i % 50;
            xs[(c * 130 + iy + 1) * 52 + ix + 1] = __float2bfloat16((float)xb[i]);
        }
        // DZ1 (429 rows x 64) -> gs[pos][oc] fp32; each thread moves 8 channels (16 bytes) at a time
        for (int i = tid; i < ROWS_PER_SAMPLE * 8; i += C1W_THREADS) {
            const int rr = i >> 3, part = i & 7, cell = part >> 1, h = part & 1;
            const int sy = rr / SG_W, sx = rr - sy * SG_W;
            const int oy = 2 * sy + (cell >> 1) - 1, ox = 2 * sx + (cell & 1) - 1;
            if (oy < 0 || oy >= 64 || ox < 0 || ox >= 25) continue;
            const uint4 v = *reinterpret_cast<const uint4*>(dz1 + ((size_t)b * ROWS_PER_SAMPLE + rr) * 64 + cell * 16 + h * 8);
            float4* d = reinterpret_cast<float4*>(gs + (oy * 25 + ox) * 16 + h * 8);
            d[0] = make_float4(bf_lo(v.x), bf_hi(v.x), bf_lo(v.y), bf_hi(v.y));
            d[1] = make_float4(bf_lo(v.z), bf_hi(v.z), bf_lo(v.w), bf_hi(v.w));
        }
        __syncthreads();
        for (int pos = warp; pos < 1600; pos += C1W_THREADS / 32) {
            const int oy = pos / 25, ox = pos - oy * 25;
            const float xv = __bfloat162float(xs[(ch * 130 + 2 * oy + ky) * 52 + 2 * ox + kx]);
            const float4* g = reinterpret_cast<const float4*>(gs + pos * 16);
#pragma unroll
            for (int qv = 0; qv < 4; ++qv) {
                const float4 gv = g[qv];
                acc[4 * qv + 0] = fmaf(xv, gv.x, acc[4 * qv + 0]);
                acc[4 * qv + 1] = fmaf(xv, gv.y, acc[4 * qv + 1]);
                acc[4 * qv + 2] = fmaf(xv, gv.z, acc[4 * qv + 2]);
                acc[4 * qv + 3] = fmaf(xv, gv.w, acc[4 * qv + 3]);
            }
        }
    }
    // cross-warp reduction through shared memory, then one atomic per weight per CTA
    __syncthreads();
    float* red = gs;                                                            // [16 warps][32 k][16 oc]
#pragma unroll
    for (int c = 0; c < 16; ++c) red[(warp * 32 + lane) * 16 + c] = acc[c];
    __syncthreads();
    for (int i = tid; i < 512; i += C1W_THREADS) {
        float s = 0.f;
        for (int w = 0; w < C1W_THREADS / 32; ++w) s += red[w * 512 + i];
        const int k = i >> 4, oc = i & 15;
        atomicAdd(&dw1[oc * 32 + k], s);                                        // conv1.weight[oc][ch][ky][kx]
    }
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

// dz2 (B*429,32), p1 (B*429,64) -> dz1 (B*429,64) bf16 (pad cells 0);  dconv1_b (16,) fp32 +=
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

// x (B,2,128,50) u8 (x_dtype 2) or f32 (0), dz1 (B*429,64) bf16 -> dconv1_w (16,2,4,4) fp32 +=
int mmg_disc_conv1_wgrad(const void* x, int x_dtype, const void* dz1, float* dconv1_w, int64_t B, void* stream) {
    MMG_REQUIRE(x && dz1 && dconv1_w && B >= 0, MMG_EINVAL, "conv1_wgrad: bad arguments");
    if (B == 0) return MMG_OK;
    const int grid = (int)(B < MMG_NUM_SMS ? B : MMG_NUM_SMS);
    if (x_dtype == 2) {
        MMG_CUDA(cudaFuncSetAttribute(conv1_wgrad_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1W_SMEM));
        conv1_wgrad_kernel<uint8_t><<<grid, C1W_THREADS, C1W_SMEM, (cudaStream_t)stream>>>((const uint8_t*)x, (const __nv_bfloat16*)dz1, dconv1_w, (int)B);
    } else if (x_dtype == 0) {
        MMG_CUDA(cudaFuncSetAttribute(conv1_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1W_SMEM));
        conv1_wgrad_kernel<float><<<grid, C1W_THREADS, C1W_SMEM, (cudaStream_t)stream>>>((const float*)x, (const __nv_bfloat16*)dz1, dconv1_w, (int)B);
    } else {
        MMG_REQUIRE(false, MMG_EINVAL, "conv1_wgrad: x_dtype must be 0 (f32) or 2 (u8)");
    }
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
