// probe kernel for mc_push_probe.py: every thread stores one 16-byte word through a multicast mapping
extern "C" __global__ void mc_store(float4 *mc, int n, float tag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float4 v = make_float4(tag, (float) i, tag + 1.0f, (float) (i * 2));
        asm volatile("multimem.st.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(mc + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}
