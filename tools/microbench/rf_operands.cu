// Register-operand bandwidth of the sm_100a fp32 datapath: cycles per instruction and SM sub-partition
// against the number of 32-bit register operands the instruction reads.  This is the measurement behind
// a packed FFMA2 with three distinct register-pair operands sustains one issue
// per THREE cycles, not two -- the register file delivers about two 32-bit operands per lane and cycle.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o rf_operands rf_operands.cu && ./rf_operands
//
// B200, 4 warps per sub-partition, 12 independent chains per thread (profiles/README.md has the output).
#include <cuda_runtime.h>
#include <cstdio>
#define ITER 2048
template <int MODE>
__global__ void __launch_bounds__(512) k(float2 *out, const float2 *in, long long *cyc)
{
    float2 a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = in[threadIdx.x + 32 * i];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int j = (i + 5) % 12, l = (i + 7) % 12;
            if (MODE == 0) a[i] = __ffma2_rn(a[j], a[l], a[i]);                        // FFMA2 R,R,R: 6 reads
            if (MODE == 1) a[i] = __fmul2_rn(a[j], a[l]);                              // FMUL2 R,R: 4 reads
            if (MODE == 2) a[i] = __ffma2_rn(a[j], make_float2(0.3f, 0.3f), a[i]);     // FFMA2 R,imm,R: 4 reads
            if (MODE == 3) a[i] = __ffma2_rn(a[j], make_float2(a[l].x, a[l].x), a[i]); // FFMA2 R,R.F32,R: 5 reads
            if (MODE == 4) { a[i].x = fmaf(a[j].x, a[l].x, a[i].x); a[i].y = fmaf(a[j].y, a[l].y, a[i].y); }   // 2 FFMA: 2 x 3 reads
            if (MODE == 5) { a[i].x = a[j].x + a[l].x; a[i].y = a[j].y + a[l].y; }    // 2 FADD: 2 x 2 reads
            if (MODE == 6) { a[i].x = fmaf(a[j].x, 0.3f, a[i].x); a[i].y = fmaf(a[j].y, 0.3f, a[i].y); }       // 2 FFMA imm: 2 x 2 reads
        }
    }
    long long t1 = clock64();
    float2 r = a[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) { r.x += a[i].x; r.y += a[i].y; }
    out[blockIdx.x * 512 + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, float2 *out, float2 *in, long long *cyc, int blocks)
{
    k<MODE><<<blocks, 512>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long *h = new long long[blocks];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < blocks; ++i) c += h[i];
    c /= blocks;
    printf("%-36s %.2f cycles per 2-wide operation per sub-partition\n", name, c / (4.0 * ITER * 12));
    delete[] h;
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount;
    float2 *in, *out;
    long long *cyc;
    cudaMalloc(&in, 8192 * sizeof(float2));
    cudaMalloc(&out, (size_t) blocks * 512 * sizeof(float2));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    cudaMemset(in, 0, 8192 * sizeof(float2));
    run<0>("FFMA2 R,R,R        (6 reads)", out, in, cyc, blocks);
    run<1>("FMUL2 R,R          (4 reads)", out, in, cyc, blocks);
    run<2>("FFMA2 R,imm,R      (4 reads)", out, in, cyc, blocks);
    run<3>("FFMA2 R,R.F32,R    (5 reads)", out, in, cyc, blocks);
    run<4>("2 x FFMA R,R,R     (2 x 3 reads)", out, in, cyc, blocks);
    run<5>("2 x FADD R,R       (2 x 2 reads)", out, in, cyc, blocks);
    run<6>("2 x FFMA R,imm,R   (2 x 2 reads)", out, in, cyc, blocks);
    return 0;
}
