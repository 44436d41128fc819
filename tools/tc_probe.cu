// tc_probe: standalone micro-experiments that pin down tcgen05 / TMA descriptor semantics on the
// B200 before the production kernels rely on them (there is no GPU in the build container).
// One generic kernel: a list of TMA 2-D box loads, then a list of tcgen05.mma with host-built
// descriptors, then a dump of TMEM.  Each experiment compares the dump with a CPU product.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I gan-des-midi-music-gen_b200/csrc tools/tc_probe.cu -o gan-des-midi-music-gen_b200/build/tc_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "tc_common.cuh"

struct LoadOp { int map, c0, c1; uint32_t smem_off, bytes; };
struct MmaOp { uint64_t a_base, b_base; uint32_t a_off, b_off, d_col, idesc, accumulate; };
struct Prog {
    int n_loads, n_mma, ncols;
    LoadOp loads[8];
    MmaOp mma[40];
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                                                    const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap m3,
                                                    const Prog* __restrict__ prog_g, float* __restrict__ out) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    __shared__ Prog prog;
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (int)(sizeof(Prog) / 4); i += 128) ((uint32_t*)&prog)[i] = ((const uint32_t*)prog_g)[i];
    if (tid == 0) { tc::mbar_init(&bar_load, 1); tc::mbar_init(&bar_mma, 1); tc::fence_barrier_init(); }
    __syncthreads();
    const int ncols = prog.ncols;
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, ncols); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const CUtensorMap* maps[4] = {&m0, &m1, &m2, &m3};
    if (tid == 0) {
        uint32_t total = 0;
        for (int i = 0; i < prog.n_loads; ++i) total += prog.loads[i].bytes;
        tc::mbar_expect_tx(&bar_load, total);
        for (int i = 0; i < prog.n_loads; ++i) {
            const LoadOp& l = prog.loads[i];
            tc::tma_load_2d(smem + l.smem_off, maps[l.map], &bar_load, l.c0, l.c1);
        }
        tc::mbar_wait(&bar_load, 0);
        tc::tc_fence_after();
        const uint32_t sbase = tc::smem_u32(smem);
        for (int i = 0; i < prog.n_mma; ++i) {
            const MmaOp& m = prog.mma[i];
            tc::mma_f16_ss(tmem + m.d_col, tc::smem_desc(m.a_base, sbase + m.a_off), tc::smem_desc(m.b_base, sbase + m.b_off), m.idesc, m.accumulate);
        }
        tc::mma_commit(&bar_mma);
    }
    __syncthreads();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tc::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(size_t)tid * ncols + c0 + j] = __uint_as_float(r[j]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------
static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }   // exact for small ints
struct Mat {                    // row-major bf16 matrix of small integers (exact in bf16)
    int rows, cols; std::vector<float> v; std::vector<uint16_t> h; void* d = nullptr;
    Mat(int r, int c, unsigned seed) : rows(r), cols(c), v((size_t)r * c), h((size_t)r * c) {
        unsigned s = seed * 2654435761u + 12345u;
        for (size_t i = 0; i < v.size(); ++i) { s = s * 1664525u + 1013904223u; v[i] = (float)((int)((s >> 16) % 7) - 3); h[i] = f2bf(v[i]); }
        cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    }
    float at(int r, int c) const { return (r >= 0 && r < rows && c >= 0 && c < cols) ? v[(size_t)r * cols + c] : 0.f; }
};

static int run(const char* name, const CUtensorMap* maps, const Prog& p, size_t smem_bytes, const std::vector<float>& want, int M, int N, int ncols,
               bool dump_layout = false) {
    Prog* dp; float* dout;
    cudaMalloc(&dp, sizeof(Prog)); cudaMemcpy(dp, &p, sizeof(Prog), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, sizeof(float) * 128 * ncols); cudaMemset(dout, 0xFF, sizeof(float) * 128 * ncols);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes + 1024);
    probe_kernel<<<1, 128, smem_bytes + 1024>>>(maps[0], maps[1], maps[2], maps[3], dp, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] CUDA ERROR: %s\n", name, cudaGetErrorString(e)); return -1; }
    std::vector<float> got((size_t)128 * ncols);
    cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double err = fabs((double)got[(size_t)m * ncols + n] - want[(size_t)m * N + n]);
        if (!(err <= 1e-3)) { if (bad < 4) printf("   [%s] mismatch m=%d n=%d got=%g want=%g\n", name, m, n, got[(size_t)m * ncols + n], want[(size_t)m * N + n]); ++bad; }
        if (err > maxerr) maxerr = err;
    }
    printf("[%s] %s  (bad=%d of %d, maxerr=%g)\n", name, bad ? "FAIL" : "PASS", bad, M * N, maxerr);
    if (dump_layout) {
        // find where logical rows landed: for each logical row m, search the lane whose first N columns match
        for (int m = 0; m < M; m += 1) {
            int found = -1;
            for (int l = 0; l < 128; ++l) { bool ok = true; for (int n = 0; n < N && ok; ++n) ok = fabs(got[(size_t)l * ncols + n] - want[(size_t)m * N + n]) < 1e-3; if (ok) { found = l; break; } }
            if (m < 4 || m % 16 == 0 || m % 16 == 15) printf("   row %d -> lane %d\n", m, found);
        }
    }
    cudaFree(dp); cudaFree(dout);
    return bad;
}

int main() {
    CUtensorMap maps[4];
    memset(maps, 0, sizeof(maps));
    const uint64_t KM128 = tc::smem_desc_base(0, 1024, tc::SW_128B);     // K-major, 128-byte rows, 8-row groups 1024 B apart
    int fails = 0;

    // ---------------- E1: plain K-major SW128 GEMM tile, M=128 N=32 K=64
    {
        Mat A(128, 64, 1), B(32, 64, 2);
        if (tc::make_map_2d_bf16(&maps[0], A.d, 64, 128, 128, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("encode failed\n"); return 2; }
        tc::make_map_2d_bf16(&maps[1], B.d, 64, 32, 128, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
        maps[2] = maps[0]; maps[3] = maps[0];
        Prog p{}; p.n_loads = 2; p.ncols = 32;
        p.loads[0] = {0, 0, 0, 0, 128 * 128}; p.loads[1] = {1, 0, 0, 16384, 32 * 128};
        p.n_mma = 4;
        for (int k = 0; k < 4; ++k) p.mma[k] = {KM128, KM128, (uint32_t)(k * 32), (uint32_t)(16384 + k * 32), 0, tc::idesc_bf16(128, 32), (uint32_t)(k > 0)};
        std::vector<float> want(128 * 32);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) { float s = 0; for (int k = 0; k < 64; ++k) s += A.at(m, k) * B.at(n, k); want[m * 32 + n] = s; }
        fails += run("E1 K-major SW128 M128 N32 K64", maps, p, 32768, want, 128, 32, 32) != 0;
        // ---------------- E4: same with M=64 -> where do the rows land?
        Prog q = p;
        for (int k = 0; k < 4; ++k) q.mma[k].idesc = tc::idesc_bf16(64, 32);
        run("E4 M=64 layout (rows 0..63)", maps, q, 32768, want, 64, 32, 32, true);
    }
    // ---------------- E2: one 144-row box, A descriptors shifted by 0/1/13/14 rows (tap-shift conv), 4 B slabs
    {
        Mat A(200, 64, 3), B(128, 64, 4);           // B: 4 taps x 32 rows
        tc::make_map_2d_bf16(&maps[0], A.d, 64, 200, 128, 64, 144, CU_TENSOR_MAP_SWIZZLE_128B);
        tc::make_map_2d_bf16(&maps[1], B.d, 64, 128, 128, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
        const int shifts[4] = {0, 1, 13, 14};
        for (int variant = 0; variant < 5; ++variant) {          // 0..3: a single shift; 4: all four accumulated
            Prog p{}; p.n_loads = 2; p.ncols = 32;
            p.loads[0] = {0, 0, 8, 0, 144 * 128};                // rows 8..151 of A
            p.loads[1] = {1, 0, 0, 18432, 128 * 128};
            std::vector<float> want(128 * 32, 0.f);
            for (int t = 0; t < 4; ++t) {
                if (variant < 4 && t != variant) continue;
                for (int k = 0; k < 4; ++k) {
                    MmaOp& m = p.mma[p.n_mma];
                    m = {KM128, KM128, (uint32_t)(shifts[t] * 128 + k * 32), (uint32_t)(18432 + t * 32 * 128 + k * 32), 0, tc::idesc_bf16(128, 32), (uint32_t)(p.n_mma > 0)};
                    ++p.n_mma;
                }
                for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) { float s = 0; for (int k = 0; k < 64; ++k) s += A.at(8 + m + shifts[t], k) * B.at(t * 32 + n, k); want[m * 32 + n] += s; }
            }
            char nm[96]; snprintf(nm, sizeof nm, variant < 4 ? "E2 row-shift %d (SW128 K-major)" : "E2 all four shifts accumulated", variant < 4 ? shifts[variant] : 0);
            fails += run(nm, maps, p, 18432 + 16384, want, 128, 32, 32) != 0;
        }
        // negative start row (TMA OOB zero fill)
        Prog p{}; p.n_loads = 2; p.ncols = 32; p.loads[0] = {0, 0, -14, 0, 144 * 128}; p.loads[1] = {1, 0, 0, 18432, 128 * 128}; p.n_mma = 4;
        for (int k = 0; k < 4; ++k) p.mma[k] = {KM128, KM128, (uint32_t)(k * 32), (uint32_t)(18432 + k * 32), 0, tc::idesc_bf16(128, 32), (uint32_t)(k > 0)};
        std::vector<float> want(128 * 32);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) { float s = 0; for (int k = 0; k < 64; ++k) s += A.at(m - 14, k) * B.at(n, k); want[m * 32 + n] = s; }
        fails += run("E2b negative TMA row coordinate -> zero fill", maps, p, 18432 + 16384, want, 128, 32, 32) != 0;
    }
    // ---------------- E6: K-major SW64 (64-byte rows, K=32 per tap), row shifts; dgrad shape M=128 N=64
    {
        Mat A(200, 32, 5), B(256, 32, 6);          // B: 4 taps x 64 rows(N) x 32 (K)
        tc::make_map_2d_bf16(&maps[0], A.d, 32, 200, 64, 32, 144, CU_TENSOR_MAP_SWIZZLE_64B);
        tc::make_map_2d_bf16(&maps[1], B.d, 32, 256, 64, 32, 256, CU_TENSOR_MAP_SWIZZLE_64B);
        const uint64_t KM64 = tc::smem_desc_base(0, 512, tc::SW_64B);
        const int shifts[4] = {14, 13, 1, 0};
        for (int variant = 0; variant < 2; ++variant) {
            Prog p{}; p.n_loads = 2; p.ncols = 64;
            p.loads[0] = {0, 0, 8, 0, 144 * 64}; p.loads[1] = {1, 0, 0, 9216, 256 * 64};
            std::vector<float> want(128 * 64, 0.f);
            for (int t = 0; t < 4; ++t) {
                if (variant == 0 && t != 1) continue;
                for (int k = 0; k < 2; ++k) {
                    MmaOp& m = p.mma[p.n_mma];
                    m = {KM64, KM64, (uint32_t)(shifts[t] * 64 + k * 32), (uint32_t)(9216 + t * 64 * 64 + k * 32), 0, tc::idesc_bf16(128, 64), (uint32_t)(p.n_mma > 0)};
                    ++p.n_mma;
                }
                for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { float s = 0; for (int k = 0; k < 32; ++k) s += A.at(8 + m + shifts[t], k) * B.at(t * 64 + n, k); want[m * 64 + n] += s; }
            }
            fails += run(variant ? "E6 SW64 K-major, 4 shifted taps, N=64" : "E6 SW64 K-major, shift 13 only", maps, p, 9216 + 16384, want, 128, 64, 64) != 0;
        }
    }
    // ---------------- E3/E5: MN-major operands (weight gradient): D[m][n] = sum_r P[r+s][m] * G[r][n]
    {
        Mat P(200, 64, 7), G(200, 32, 8);
        tc::make_map_2d_bf16(&maps[0], P.d, 64, 200, 128, 64, 96, CU_TENSOR_MAP_SWIZZLE_128B);   // 96 rows x 64 ch box
        tc::make_map_2d_bf16(&maps[1], G.d, 32, 200, 64, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B);     // 64 rows x 32 oc box
        // A: MN-major SW128, M = 128 = two 64-channel atoms: atom1 = the same box shifted by one row (LBO = 128 B)
        // K = rows: 8-row groups are 1024 B apart (SBO)
        const int shifts[3] = {0, 1, 13};
        for (int variant = 0; variant < 4; ++variant) {
            const bool m64 = variant == 3;
            const int s = m64 ? 13 : shifts[variant];
            const uint64_t A_MN = tc::smem_desc_base(128, 1024, tc::SW_128B);
            const uint64_t B_MN = tc::smem_desc_base(0, 512, tc::SW_64B);
            Prog p{}; p.n_loads = 2; p.ncols = 32;
            p.loads[0] = {0, 0, 16, 0, 96 * 128}; p.loads[1] = {1, 0, 16, 12288, 64 * 64};
            p.n_mma = 4;
            const int M = m64 ? 64 : 128;
            for (int k = 0; k < 4; ++k)
                p.mma[k] = {A_MN, B_MN, (uint32_t)(s * 128 + k * 16 * 128), (uint32_t)(12288 + k * 16 * 64), 0, tc::idesc_bf16(M, 32, 1, 1), (uint32_t)(k > 0)};
            std::vector<float> want(128 * 32);
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
                float acc = 0;
                for (int r = 0; r < 64; ++r) acc += P.at(16 + r + s + (m >> 6), m & 63) * G.at(16 + r, n);
                want[m * 32 + n] = acc;
            }
            char nm[96]; snprintf(nm, sizeof nm, "E3 MN-major A(SW128,LBO=1 row)/B(SW64) M=%d shift %d", M, s);
            int r = run(nm, maps, p, 12288 + 4096, want, M, 32, 32, m64);
            if (!m64) fails += r != 0;
        }
    }

    // ---------------- E7: conv1 forward shape.  A = 16-byte rows (8 values per super pixel), K=16 = two OVERLAPPING rows
    //                  (no swizzle, K-major: LBO = 16 B between the two K chunks, SBO = 128 B between 8-row groups)
    {
        Mat A(400, 8, 9), Bw(32, 8, 10);       // Bw flat rows: ((ty*2 + c)*2 + g)*8 + i  <->  n = g*8+i, k = ty*16 + c*8 + e
        CUtensorMap mA, mB;
        {
            auto fn = tc::get_encode_fn();
            cuuint64_t dims[2] = {8, 400}; cuuint64_t strides[1] = {16}; cuuint32_t box[2] = {8, 160}; cuuint32_t es[2] = {1, 1};
            fn(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A.d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint64_t dimsb[2] = {8, 32}; cuuint32_t boxb[2] = {8, 32};
            fn(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Bw.d, dimsb, strides, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        maps[0] = mA; maps[1] = mB;
        std::vector<float> want(128 * 16);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
            float acc = 0;
            for (int ty = 0; ty < 2; ++ty) for (int c = 0; c < 2; ++c) for (int e = 0; e < 8; ++e)
                acc += A.at(40 + m + ty * 26 + c, e) * Bw.at(((ty * 2 + c) * 2 + (n >> 3)) * 8 + (n & 7), e);
            want[m * 16 + n] = acc;
        }
        // which of LBO / SBO is the K-direction stride for un-swizzled K-major operands?  try all four assignments
        for (int va = 0; va < 2; ++va) for (int vb = 0; vb < 2; ++vb) {
            const uint64_t A_K = va ? tc::smem_desc_base(128, 16, tc::SW_NONE) : tc::smem_desc_base(16, 128, tc::SW_NONE);
            const uint64_t B_K = vb ? tc::smem_desc_base(128, 256, tc::SW_NONE) : tc::smem_desc_base(256, 128, tc::SW_NONE);
            Prog p{}; p.n_loads = 2; p.ncols = 32;
            p.loads[0] = {0, 0, 40, 0, 160 * 16}; p.loads[1] = {1, 0, 0, 4096, 32 * 16};
            p.n_mma = 2;
            for (int ty = 0; ty < 2; ++ty) p.mma[ty] = {A_K, B_K, (uint32_t)(ty * 26 * 16), (uint32_t)(4096 + ty * 512), 0, tc::idesc_bf16(128, 16), (uint32_t)(ty > 0)};
            char nm[128]; snprintf(nm, sizeof nm, "E7 conv1-fwd no-swizzle K-major: A(K-stride in %s) B(K-stride in %s)", va ? "SBO" : "LBO", vb ? "SBO" : "LBO");
            run(nm, maps, p, 8192, want, 128, 16, 32);
        }
        // B as a plain SW32 K-major operand [16 n][16 k] per ty (32-byte rows), A as above in both assignments
        {
            Mat B2(32, 16, 12);                  // rows ty*16 + n, 16 k values: k = c*8 + e
            CUtensorMap mB2;
            tc::make_map_2d_bf16(&mB2, B2.d, 16, 32, 32, 16, 32, CU_TENSOR_MAP_SWIZZLE_32B);
            maps[1] = mB2;
            std::vector<float> want3(128 * 16);
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
                float acc = 0;
                for (int ty = 0; ty < 2; ++ty) for (int c = 0; c < 2; ++c) for (int e = 0; e < 8; ++e)
                    acc += A.at(40 + m + ty * 26 + c, e) * B2.at(ty * 16 + n, c * 8 + e);
                want3[m * 16 + n] = acc;
            }
            for (int va = 0; va < 2; ++va) {
                const uint64_t A_K = va ? tc::smem_desc_base(128, 16, tc::SW_NONE) : tc::smem_desc_base(16, 128, tc::SW_NONE);
                const uint64_t B_K = tc::smem_desc_base(0, 256, tc::SW_32B);
                Prog p{}; p.n_loads = 2; p.ncols = 32;
                p.loads[0] = {0, 0, 40, 0, 160 * 16}; p.loads[1] = {1, 0, 0, 4096, 32 * 32};
                p.n_mma = 2;
                for (int ty = 0; ty < 2; ++ty) p.mma[ty] = {A_K, B_K, (uint32_t)(ty * 26 * 16), (uint32_t)(4096 + ty * 512), 0, tc::idesc_bf16(128, 16), (uint32_t)(ty > 0)};
                char nm[128]; snprintf(nm, sizeof nm, "E7b conv1-fwd: A no-swizzle (K-stride in %s), B SW32 K-major", va ? "SBO" : "LBO");
                run(nm, maps, p, 8192, want3, 128, 16, 32);
            }
            maps[1] = mB;
        }

        // ---------------- E8: conv1 wgrad shape.  A = same rows, MN-major no-swizzle: M = 64 = 8 atoms of 8 values, atom j = row + j (SBO = 16 B),
        //                  K rows 16 B apart, 8-row groups 128 B apart (LBO);  B = G rows of 16 oc (32 B), MN-major SW32 (SBO = 256 B)
        Mat G(400, 16, 11);
        CUtensorMap mG;
        tc::make_map_2d_bf16(&mG, G.d, 16, 400, 32, 16, 64, CU_TENSOR_MAP_SWIZZLE_32B);
        maps[1] = mG;
        const uint64_t A_MN = tc::smem_desc_base(128, 16, tc::SW_NONE), B_MN = tc::smem_desc_base(0, 256, tc::SW_32B);
        Prog q{}; q.n_loads = 2; q.ncols = 32;
        q.loads[0] = {0, 0, 40, 0, 160 * 16}; q.loads[1] = {1, 0, 40, 4096, 64 * 32};
        q.n_mma = 4;
        for (int k = 0; k < 4; ++k) q.mma[k] = {A_MN, B_MN, (uint32_t)(k * 16 * 16), (uint32_t)(4096 + k * 16 * 32), 0, tc::idesc_bf16(64, 16, 1, 1), (uint32_t)(k > 0)};
        std::vector<float> want2(64 * 16);
        for (int m = 0; m < 64; ++m) for (int n = 0; n < 16; ++n) {
            float acc = 0;
            for (int r = 0; r < 64; ++r) acc += A.at(40 + r + (m >> 3), m & 7) * G.at(40 + r, n);
            want2[m * 16 + n] = acc;
        }
        // M=64 results live in lanes (m%16) + 32*(m/16): compare through the lane map
        {
            Prog* dp; float* dout;
            cudaMalloc(&dp, sizeof(Prog)); cudaMemcpy(dp, &q, sizeof(Prog), cudaMemcpyHostToDevice);
            cudaMalloc(&dout, sizeof(float) * 128 * 32); cudaMemset(dout, 0, sizeof(float) * 128 * 32);
            probe_kernel<<<1, 128, 8192 + 1024>>>(maps[0], maps[1], maps[2], maps[3], dp, dout);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> got(128 * 32);
            cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 64; ++m) for (int n = 0; n < 16; ++n) {
                const int lane = (m % 16) + 32 * (m / 16);
                if (fabs(got[lane * 32 + n] - want2[m * 16 + n]) > 1e-3) { if (bad < 4) printf("   [E8] m=%d n=%d got=%g want=%g\n", m, n, got[lane * 32 + n], want2[m * 16 + n]); ++bad; }
            }
            printf("[E8 conv1-wgrad: MN-major no-swizzle A (atoms 1 row apart) x MN-major SW32 B, M=64 N=16] %s (err=%s bad=%d)\n", bad || e ? "FAIL" : "PASS", cudaGetErrorString(e), bad);
            fails += (bad || e) ? 1 : 0;
        }
    }
    printf("tc_probe: %d failing experiment(s)\n", fails);
    return 0;
}
