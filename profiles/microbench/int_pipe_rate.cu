// profiles/microbench/int_pipe_rate.cu -- measured issue rate of the integer/DPX instructions the alignment kernels are
// built from, in thread-instructions per clock per SM (B200, sm_100a).  This is the denominator of the int-pipe roofline
// (SURVEY.md 8d asks for "measured peak int32/DPX issue rate"); MEASURED_PEAKS.json only carries HBM and bf16 numbers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe_rate int_pipe_rate.cu && ./int_pipe_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
    if (OP == 0) return a + b;                                  // IADD3
    if (OP == 1) return __viaddmax_s16x2(a, b, c);              // VIADDMNMX.S16x2
    if (OP == 2) return __viaddmax_s16x2_relu(a, b, c);
    if (OP == 3) return __viaddmin_s16x2_relu(a, b, c);
    if (OP == 4) return __vimax3_s16x2(a, b, c);                // VIMNMX3.S16x2
    if (OP == 5) return __vmaxs2(a, b);                         // VIMNMX.S16x2
    if (OP == 6) return (uint32_t) __viaddmax_s32((int) a, (int) b, (int) c);
    if (OP == 7) return (uint32_t) __vimax3_s32((int) a, (int) b, (int) c);
    if (OP == 8) return __byte_perm(a, b, 0x5140 ^ (c & 0));   // PRMT
    if (OP == 9) return a * b + c;                              // IMAD (fma pipe)
    if (OP == 10) return (a & b) ^ c;                           // LOP3
    if (OP == 11) return (uint32_t) max((int) a, (int) (b ^ c)) + 1u;   // VIMNMX.S32 + one add (a bare max(a,b) chain reaches a fixed point and is folded away)
    if (OP == 12) return __vadd2(a, b);                         // packed add (may be emulated)
    if (OP == 13) return __funnelshift_l(a, b, 16);             // SHF
    return a;
}

template <int OP>
__global__ void __launch_bounds__(1024) rate_kernel(uint32_t *out, long long *cycles, uint32_t seed) {
    uint32_t x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = seed * (threadIdx.x + 1) + i; y[i] = (seed >> 2) * (threadIdx.x + 3) + 7 * i; }
    const uint32_t c = seed >> 3;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = op<OP>(x[i], y[i], c);   // x and y feed each other: nothing folds
#pragma unroll
            for (int i = 0; i < ILP; i++) y[i] = op<OP>(y[i], x[i], c);
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// mixed: per iteration 2 DPX + 1 IMAD, to see whether the fma pipe issues alongside the alu pipe
__global__ void __launch_bounds__(1024) mix_kernel(uint32_t *out, long long *cycles, uint32_t seed) {
    uint32_t x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = seed * (threadIdx.x + 1) + i; y[i] = x[i] ^ 77u; }
    const uint32_t b = seed | 1u, c = seed >> 3;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                x[i] = __viaddmax_s16x2(x[i], b, c);
                y[i] = y[i] * b + c;
                x[i] = __vimax3_s16x2(x[i], b, c);
            }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// shared-memory LDS.128 rate, conflict-free
__global__ void __launch_bounds__(1024) lds_kernel(uint32_t *out, long long *cycles, uint32_t seed) {
    __shared__ uint4 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_uint4(i, seed, i ^ seed, 1);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    int idx = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint4 v = buf[(idx + r * 32) & 2047];
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        idx += 256;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x ^ acc.y ^ acc.z ^ acc.w;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(1024) shfl_kernel(uint32_t *out, long long *cycles, uint32_t seed) {
    uint32_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = seed * (threadIdx.x + 1) + i;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char *name, F launch, double ops_per_thread, uint32_t *d_out, long long *d_cyc, int blocks) {
    launch(); launch();
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[1024];
    cudaMemcpy(cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; i++) mean += (double) cyc[i];
    mean /= blocks;
    const double per_clk_sm = ops_per_thread * 1024.0 / mean;
    printf("%-28s %8.2f thread-instr/clk/SM   (%.0f cycles, %.3f ms, implied clock %.0f MHz)\n", name, per_clk_sm, mean, ms,
           mean / (ms * 1e3));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount;  // one 1024-thread CTA per SM
    printf("device %s, %d SMs\n", p.name, blocks);
    uint32_t *d_out; long long *d_cyc;
    cudaMalloc(&d_out, sizeof(uint32_t) * 1024 * blocks);
    cudaMalloc(&d_cyc, sizeof(long long) * blocks);
    const double n = (double) ITERS * 4 * ILP;
#define RUN(OP, NAME) run(NAME, [&] { rate_kernel<OP><<<blocks, 1024>>>(d_out, d_cyc, 12345u); }, n, d_out, d_cyc, blocks)
    RUN(0, "IADD");
    RUN(1, "viaddmax_s16x2");
    RUN(2, "viaddmax_s16x2_relu");
    RUN(3, "viaddmin_s16x2_relu");
    RUN(4, "vimax3_s16x2");
    RUN(5, "vmaxs2");
    RUN(6, "viaddmax_s32");
    RUN(7, "vimax3_s32");
    RUN(8, "PRMT");
    RUN(9, "IMAD");
    RUN(10, "LOP3");
    run("max_s32+IADD (2 ops)", [&] { rate_kernel<11><<<blocks, 1024>>>(d_out, d_cyc, 12345u); }, 2 * n, d_out, d_cyc, blocks);
    RUN(12, "vadd2");
    RUN(13, "SHF funnelshift");
    run("mix 2xDPX+1xIMAD (3 ops)", [&] { mix_kernel<<<blocks, 1024>>>(d_out, d_cyc, 12345u); }, (double) ITERS * 2 * ILP * 3, d_out, d_cyc, blocks);
    run("LDS.128 (per 16B load)", [&] { lds_kernel<<<blocks, 1024>>>(d_out, d_cyc, 12345u); }, (double) ITERS * 8, d_out, d_cyc, blocks);
    run("SHFL.UP", [&] { shfl_kernel<<<blocks, 1024>>>(d_out, d_cyc, 12345u); }, (double) ITERS * ILP, d_out, d_cyc, blocks);
    return 0;
}
