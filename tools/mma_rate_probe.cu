// Issue-rate probe for the tcgen05.mma shapes / operand formats of the discriminator pass kernel (csrc/disc_tc_pass.cu):
// one CTA per SM, one thread issues NREP MMAs of a given configuration back to back (operands = whatever shared memory holds,
// all finite), commits, waits; SM clock before the first issue and after the commit's arrival -> cycles per MMA.
// Varied: operand majors / swizzles / N as used by conv1, conv2, conv2 wgrad, conv2 dgrad, conv1 wgrad; accumulating into ONE
// TMEM accumulator (a dependent chain, as a K loop does) or rotating over 2 / 4 accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gan-des-midi-music-gen_b200/csrc tools/mma_rate_probe.cu -o tools/mma_rate_probe.bin
#include <cstdio>
#include "tc_common.cuh"

#include <cstdlib>

template <int A_STEP, int B_STEP, int N_COLS, int N_ACC>
__global__ void __launch_bounds__(128, 1) probe(uint64_t da, uint64_t db, uint32_t idesc, long long* out) {
    constexpr int nrep = 64;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 7);      // finite bf16 pairs
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_s, 512); tc::tmem_relinquish(); }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;
    if (threadIdx.x < 32 && tc::elect_one()) {
        const uint32_t a0 = tc::smem_u32(smem), b0 = tc::smem_u32(smem + 128 * 1024);
        for (int round = 0; round < 2; ++round) {                     // round 0 warms up
            const long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < nrep; ++i)          // fully unrolled: descriptors are uniform-register constants, the UTCHMMA stream is back to back
                tc::mma_f16_ss(tmem + (uint32_t)((i % N_ACC) * N_COLS), tc::smem_desc(da, a0 + (uint32_t)(i % 16) * A_STEP), tc::smem_desc(db, b0 + (uint32_t)(i % 8) * B_STEP), idesc,
                               i >= N_ACC ? 1u : 0u);
            tc::mma_commit(&bar);
            tc::mbar_wait(&bar, (uint32_t)round);
            const long long t1 = clock64();
            if (round == 1 && blockIdx.x == 0) out[0] = (t1 - t0) * 4;      // (x4: the host divides by 256)
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

template <int A_STEP, int B_STEP, int N_COLS>
void run(const char* name, uint64_t da, uint64_t db, uint32_t idesc, long long* out) {
    printf("%-100s", name);
    auto one = [&](auto kern, int n_acc) {
        if (n_acc * N_COLS > 512) { printf(" %10s", "-"); return; }
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
        kern<<<148, 128, 201 * 1024>>>(da, db, idesc, out);
        long long cyc = 0;
        cudaError_t e = cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf(" err %s\n", cudaGetErrorString(e)); exit(1); }
        printf(" %10.1f", (double)cyc / 256);
    };
    one(probe<A_STEP, B_STEP, N_COLS, 1>, 1);
    one(probe<A_STEP, B_STEP, N_COLS, 2>, 2);
    one(probe<A_STEP, B_STEP, N_COLS, 4>, 4);
    printf("\n");
}

int main() {
    using namespace tc;
    long long* out;
    cudaMalloc(&out, 8);
    printf("%-100s %10s %10s %10s\n", "configuration (cycles per MMA, 64 MMAs back to back, fully unrolled issue)", "1 acc", "2 accs", "4 accs");
    run<2048, 512, 16>("conv1 fwd   A K-major SW_NONE (16 B rows), B K-major SW32,  N=16", smem_desc_base(16, 128, SW_NONE), smem_desc_base(0, 256, SW_32B), idesc_bf16(128, 16), out);
    run<32, 32, 32>("conv2 fwd   A K-major SW128, B K-major SW128, N=32 (K slices of the same rows)", smem_desc_base(0, 1024, SW_128B), smem_desc_base(0, 1024, SW_128B), idesc_bf16(128, 32), out);
    run<4096, 32, 32>("conv2 fwd   same, A advancing by 32 rows per MMA", smem_desc_base(0, 1024, SW_128B), smem_desc_base(0, 1024, SW_128B), idesc_bf16(128, 32), out);
    run<32, 32, 64>("conv2 dgrad A K-major SW64, B K-major SW64, N=64", smem_desc_base(0, 512, SW_64B), smem_desc_base(0, 512, SW_64B), idesc_bf16(128, 64), out);
    run<2048, 1024, 64>("conv2 wgrad A MN-major SW128 (2 atoms 128 B apart), B MN-major SW64 (2 atoms 832 B apart), N=64", smem_desc_base(128, 1024, SW_128B), smem_desc_base(832, 512, SW_64B),
                        idesc_bf16(128, 64, 1, 1), out);
    run<2048, 1024, 32>("conv2 wgrad (round 1) same A, B MN-major SW64 one atom, N=32", smem_desc_base(128, 1024, SW_128B), smem_desc_base(0, 512, SW_64B), idesc_bf16(128, 32, 1, 1), out);
    run<256, 2048, 64>("conv1 wgrad A MN-major SW_NONE (16 atoms 6912 B apart), B MN-major SW128 one atom, N=64", smem_desc_base(128, 6912, SW_NONE), smem_desc_base(128, 1024, SW_128B),
                       idesc_bf16(128, 64, 1, 1), out);
    run<32, 32, 256>("reference   A K-major SW128, B K-major SW128, N=256", smem_desc_base(0, 1024, SW_128B), smem_desc_base(0, 1024, SW_128B), idesc_bf16(128, 256), out);
    run<32, 32, 128>("reference   A K-major SW128, B K-major SW128, N=128", smem_desc_base(0, 1024, SW_128B), smem_desc_base(0, 1024, SW_128B), idesc_bf16(128, 128), out);
    run<32, 32, 64>("reference   A K-major SW128, B K-major SW128, N=64", smem_desc_base(0, 1024, SW_128B), smem_desc_base(0, 1024, SW_128B), idesc_bf16(128, 64), out);
    run<32, 32, 16>("reference   A K-major SW128, B K-major SW128, N=16", smem_desc_base(0, 1024, SW_128B), smem_desc_base(0, 1024, SW_128B), idesc_bf16(128, 16), out);
    return 0;
}
