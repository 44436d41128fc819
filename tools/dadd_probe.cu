// Dependent fp64 add latency on the GPU and the cost of the rasteriser's chain loop variants (the floor of its running time sum).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dadd_probe.bin tools/dadd_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 512;
__global__ void pure(double* out, const double* in, int n, long long* cyc) {
    double t = in[0];
    const double d = in[1];
    long long c0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) t = __dadd_rn(t, d);
    long long c1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = c1 - c0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// variant: 0 = LDS.64 operand, no store; 1 = LDS.64 operand + STS.64 result in place; 2 = the K1 loop (LDS.128 prefetch one trip ahead, STS.128 in place);
// 3 = K1 loop without the stores; 4 = LDS.128 operands for the whole chunk into registers first (64 doubles), then DADDs, then STS.128
template <int V>
__global__ void loop(double* out, const double* in, int chunks, long long* cyc) {
    __shared__ __align__(16) double tbuf[4][CH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* tb = tbuf[warp & 3];
    for (int i = lane; i < CH; i += 32) tb[i] = in[1];
    __syncwarp();
    double t = in[0];
    long long c0 = clock64();
    if (lane == 0) {
        for (int c = 0; c < chunks; ++c) {
            if (V == 0) {
#pragma unroll 8
                for (int k = 0; k < CH; ++k) t = __dadd_rn(t, tb[k]);
            } else if (V == 1) {
#pragma unroll 8
                for (int k = 0; k < CH; ++k) { t = __dadd_rn(t, tb[k]); tb[k] = in[1] == 0.5 ? t : 0.0123; }
            } else if (V == 2 || V == 3) {
                double2* tb2 = reinterpret_cast<double2*>(tb);
                double2 v[4], w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = tb2[u];
                for (int k = 0; k < CH; k += 8) {
                    if (k + 8 < CH) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) w[u] = tb2[(k >> 1) + 4 + u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        t = __dadd_rn(t, v[u].x); v[u].x = in[1] == 0.5 ? t : 0.0123;
                        t = __dadd_rn(t, v[u].y); v[u].y = in[1] == 0.5 ? t : 0.0123;
                        if (V == 2) tb2[(k >> 1) + u] = v[u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = w[u];
                }
            } else {
                double2* tb2 = reinterpret_cast<double2*>(tb);
                for (int k0 = 0; k0 < CH; k0 += 64) {
                    double2 v[32];
#pragma unroll
                    for (int u = 0; u < 32; ++u) v[u] = tb2[(k0 >> 1) + u];
#pragma unroll
                    for (int u = 0; u < 32; ++u) {
                        t = __dadd_rn(t, v[u].x); v[u].x = in[1] == 0.5 ? t : 0.0123;
                        t = __dadd_rn(t, v[u].y); v[u].y = in[1] == 0.5 ? t : 0.0123;
                    }
#pragma unroll
                    for (int u = 0; u < 32; ++u) tb2[(k0 >> 1) + u] = v[u];
                }
            }
        }
    }
    long long c1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = c1 - c0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = t + tb[lane];
}
// variant 5: K1 loop with the prefix sums rounded on the spot and stored as packed u16 steps (one STS.128 per 8 elements)
__global__ void loop_steps(double* out, const double* in, int chunks, long long* cyc) {
    __shared__ __align__(16) double tbuf[4][CH];
    __shared__ __align__(16) unsigned short sbuf[4][CH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* tb = tbuf[warp & 3];
    uint4* sb = reinterpret_cast<uint4*>(sbuf[warp & 3]);
    for (int i = lane; i < CH; i += 32) tb[i] = in[1];
    __syncwarp();
    double t = in[0];
    long long c0 = clock64();
    if (lane == 0) {
        for (int c = 0; c < chunks; ++c) {
            const double2* tb2 = reinterpret_cast<const double2*>(tb);
            double2 v[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = tb2[u];
            for (int k = 0; k < CH; k += 8) {
                if (k + 8 < CH) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) w[u] = tb2[(k >> 1) + 4 + u];
                }
                unsigned s[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    t = __dadd_rn(t, v[u].x); s[2 * u] = min((unsigned)__double2int_rn(t), 65535u);
                    t = __dadd_rn(t, v[u].y); s[2 * u + 1] = min((unsigned)__double2int_rn(t), 65535u);
                }
                sb[k >> 3] = make_uint4(s[0] | (s[1] << 16), s[2] | (s[3] << 16), s[4] | (s[5] << 16), s[6] | (s[7] << 16));
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = w[u];
            }
        }
    }
    long long c1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = c1 - c0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = t + tb[lane] + sbuf[warp & 3][lane];
}
template <int V>
void run(const char* name, double* out, double* in, long long* cyc) {
    const int chunks = 64;
    for (int warps : {1, 4, 12}) {
        loop<V><<<148, warps * 32>>>(out, in, chunks, cyc);
        cudaDeviceSynchronize();
        loop<V><<<148, warps * 32>>>(out, in, chunks, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-58s warps/SM %2d: %.2f cycles per element\n", name, warps, (double)c / (chunks * CH));
    }
}
int main() {
    double *in, *out; long long* cyc;
    cudaMalloc(&in, 16); cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
    double h[2] = {0.0, 0.0123};
    cudaMemcpy(in, h, 16, cudaMemcpyHostToDevice);
    const int n = 1 << 16;
    for (int warps : {1, 12, 32}) {
        pure<<<148, warps * 32>>>(out, in, n, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-58s warps/SM %2d: %.2f cycles per element\n", "pure dependent DADD chain", warps, (double)c / n);
    }
    run<0>("LDS.64 operand, no store", out, in, cyc);
    run<1>("LDS.64 operand + STS.64 result in place", out, in, cyc);
    run<3>("K1 loop without the stores (LDS.128 one trip ahead)", out, in, cyc);
    run<2>("K1 loop (LDS.128 one trip ahead, STS.128 in place)", out, in, cyc);
    run<4>("64 operands into registers, 64 DADDs, 32 STS.128", out, in, cyc);
    for (int warps : {1, 4, 12}) {
        loop_steps<<<148, warps * 32>>>(out, in, 64, cyc);
        cudaDeviceSynchronize();
        loop_steps<<<148, warps * 32>>>(out, in, 64, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-58s warps/SM %2d: %.2f cycles per element\n", "K1 loop, rounded on the spot, u16 steps (1 STS.128 / 8)", warps, (double)c / (64 * CH));
    }
    return 0;
}
