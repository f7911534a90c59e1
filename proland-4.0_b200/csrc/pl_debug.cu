/*
 * pl_debug.cu -- test hooks (exported, prefixed pl_debug_): evaluate the
 * branch-free fp32 helpers of pl_fpexact.cuh next to the plain IEEE operators
 * on the device, so the tests can assert bit-equality on the GPU itself.
 */
#include "pl_internal.h"
#include "pl_fpexact.cuh"

namespace {
__global__ void fpexact_kernel(int n, const float *a, const float *b, float *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = a[i], y = b[i];
    const float ry = plfp::rcp_rn(y);
    out[i] = plfp::div_rn(x, y, ry);
    out[n + i] = x / y;
    out[2 * n + i] = ry;
    out[3 * n + i] = 1.0f / y;
    out[4 * n + i] = plfp::sqrt_rn(fabsf(x));
    out[5 * n + i] = sqrtf(fabsf(x));
}
}  // namespace

/* a, b: n host floats; out: 6*n host floats = (div_rn, a/b, rcp_rn, 1/b, sqrt_rn, sqrtf)(|a|) */
extern "C" int pl_debug_fpexact(pl_ctx *ctx, int n, const float *a, const float *b, float *out)
{
    if (!ctx || !a || !b || !out || n <= 0) return pl_set_error(PL_ERR_ARG, "bad argument");
    PL_CUDA(cudaSetDevice(ctx->device));
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    PL_CUDA(cudaMalloc(&da, sizeof(float) * n));
    PL_CUDA(cudaMalloc(&db, sizeof(float) * n));
    PL_CUDA(cudaMalloc(&dout, sizeof(float) * 6 * (size_t) n));
    PL_CUDA(cudaMemcpyAsync(da, a, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    PL_CUDA(cudaMemcpyAsync(db, b, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    fpexact_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, da, db, dout);
    PL_CUDA(cudaGetLastError());
    PL_CUDA(cudaMemcpyAsync(out, dout, sizeof(float) * 6 * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return PL_OK;
}
